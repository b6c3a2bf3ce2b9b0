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

CFG = dict(W=80, H=80, B=64, nfeature=64, nlayer=4, code=256, n_lis=1, lr=2e-5, lambda_r=0.9)
WORKLOAD = "G-LIS 1 LIS module, 80x80 CelebA-shaped synthetic, batch 64/GPU (BASELINE configs[1])"
# F_step = 8 F_D + 4 F_G + 4 F_LIS = 250.9 GFLOP at config 2 (B = 64)
GFLOP_PER_STEP = 250.9
SEED = 1234


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
                                          CFG["n_lis"], "fractional")
    d = oracle.build_discriminator(CFG["W"], CFG["H"], CFG["nfeature"], CFG["nlayer"], "weight", 0)
    return g, d


def time_cpu_oracle(steps, warmup, budget_s=150.0):
    """(seconds per step, cores, images per step) of the oracle's iteration on all host cores.

    A step is one full config-2 iteration at batch 64; if `steps + warmup` of those would not fit
    `budget_s`, the per-step batch is cut (never below 8) so that the run stays bounded — images/s
    is then per-step images over per-step time, which is what the metric means."""
    from oracle.step import GLISOracleTrainer
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g, d = build_oracle_pair()
    tr = GLISOracleTrainer(g, d, lr=CFG["lr"], lambda_r=CFG["lambda_r"])
    gen = torch.Generator().manual_seed(SEED + 1)

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


def run_reference(args, rank):
    if rank != 0:
        return
    sec, cores, b_step = time_cpu_oracle(args.steps, args.warmup)
    ips = b_step / sec
    sample = "%d full config-2 iterations (batch %d per step) after %d warm-up, oracle port on torch %s CPU, %s" % (
        args.steps, b_step, args.warmup, torch.__version__, cpu_model())
    print(json.dumps({
        "impl": "reference", "metric": "G-LIS train images/sec at 80x80 bs64/GPU", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": b_step,
                   "note": "python2.7/PyTorch@065c5986 pin not installable offline; nearest installable "
                           "PyTorch CPU build used (oracle port of the reference step)"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def probe_dominant_kernel(dev, precision_name):
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
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get("tc_conv_kernel_d1_2B_dram_bytes")
    return {"bound": "tensor", "achieved": flop / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "traffic": traffic,
            "kernel": "tc_conv_kernel  D level-1 conv 64->128, 40x40->20x20, 128 images (M=51200 N=128 K=1024)",
            "ms_per_launch": ms, "algorithmic_gflop_per_launch": flop / 1e9,
            "mma_passes": passes, "mma_pipe_tflops": passes * flop / (ms * 1e-3) / 1e12,
            "note": "bf16x3 issues 3 MMAs per algorithmic product (fp32-faithful split); "
                    "`achieved` counts algorithmic FLOPs only"}


def run_ours(args, rank, world, local):
    import torch.distributed as dist
    import common.model as pm
    from glis_b200 import _lib, dp, ops
    from glis_b200.trainer import GLISTrainer, GraphedStep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: the product path needs a CUDA device (there is no CPU fallback)")
    if args.precision:
        _lib.set_precision(args.precision)
    prec_name = [k for k, v in _lib.PRECISION_NAMES.items() if v == _lib.default_precision][0]
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    data_seed = dp.seed_everything(SEED, rank)
    gen = pm.GeneratorLearnedInputSpace(CFG["W"], CFG["H"], CFG["nfeature"], CFG["nlayer"], CFG["code"], "weight",
                                        CFG["n_lis"], "fractional").to(dev)
    dis = pm.build_discriminator(CFG["W"], CFG["H"], CFG["nfeature"], CFG["nlayer"], "weight", 0).to(dev)
    sync = dp.OverlappedGradSync(world) if world > 1 else None
    tr = GLISTrainer(gen, dis, lr=CFG["lr"], lambda_r=CFG["lambda_r"], grad_sync=sync)
    B, H, W, code = CFG["B"], CFG["H"], CFG["W"], CFG["code"]
    depth = CFG["n_lis"]  # LIS depth forced to "all": fixed work per step (the stochastic schedule halves LIS work)

    # ---- device-resident inputs (value): Philox on the device, a fresh batch every step
    real = torch.empty(B, 3, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    zd, zg = torch.empty(B, code, device=dev), torch.empty(B, code, device=dev)
    counter = [0]

    graphed = None
    if not args.no_graph:
        graphed = GraphedStep(tr, B, H, W, code, dev)
        real, zd, zg = graphed.real, graphed.z_d, graphed.z_g   # fill the static buffers in place

    def device_step():
        off = counter[0] * (1 << 22)
        counter[0] += 1
        ops.uniform_(real, data_seed, off)
        ops.randn_(zd, data_seed + 7, off)
        ops.randn_(zg, data_seed + 13, off)
        if graphed is not None:
            return graphed.step(None, None, None, depth, depth)
        return tr.step(real, zd, zg, depth, depth)

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
        device_step()
    launches0 = _lib.launch_count
    ms_total = timed_region(device_step, args.steps)
    launches = _lib.launch_count - launches0
    if graphed is not None:
        # a replayed graph launches the kernels captured once: count them from an eager step
        l0 = _lib.launch_count
        tr.step(real, zd, zg, depth, depth)
        launches = args.steps * (_lib.launch_count - l0 + 3)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = B * world / (ms_step * 1e-3)

    # ---- per-kernel timing for the roofline (separate short pass; events serialise nothing)
    _lib.timed.enabled = True
    for _ in range(min(args.steps, 5)):
        ops.uniform_(real, data_seed, 1 << 40)
        tr.step(real, zd, zg, depth, depth)      # eager: the event brackets cannot live inside a graph
    torch.cuda.synchronize()
    _lib.timed.enabled = False
    per_kernel = _lib.timer_summary()
    if args.kernel_table and rank == 0:
        with open(args.kernel_table, "w") as fh:
            json.dump({k: {"launches": c, "ms": m} for k, (c, m) in sorted(per_kernel.items())}, fh, indent=1)

    probe = probe_dominant_kernel(dev, prec_name) if rank == 0 else None

    # ---- end to end: pinned host buffers -> H2D -> step -> D2H of the losses, through the public
    # host-fed API (trainer.HostFedStepper): every iteration's image + noise batches cross PCIe from
    # pinned memory and every iteration's losses are read back on the host; the copy of iteration
    # i+1 overlaps the compute of iteration i and the host reads losses one iteration late.
    h_real = torch.rand(B, 3, H, W).pin_memory()
    h_zd, h_zg = torch.randn(B, code).pin_memory(), torch.randn(B, code).pin_memory()
    if graphed is not None:
        from glis_b200.trainer import HostFedStepper
        feeder = HostFedStepper(graphed)
        h2d, d2h = feeder.h2d_bytes, 4 * (3 + CFG["n_lis"])
        last = [None]

        def e2e_step():
            last[0] = feeder.submit(h_real, h_zd, h_zg, depth, depth)

        def e2e_drain():
            last[0] = feeder.flush()
    else:
        d_real = torch.empty(B, 3, H, W, device=dev)
        h_loss = torch.empty(3).pin_memory()
        h2d = h_real.numel() * 4 + h_zd.numel() * 4 + h_zg.numel() * 4
        d2h = h_loss.numel() * 4

        def e2e_step():
            d_real.copy_(h_real, non_blocking=True)
            zd.copy_(h_zd, non_blocking=True)
            zg.copy_(h_zg, non_blocking=True)
            out = tr.step(d_real, zd, zg, depth, depth)
            h_loss.copy_(torch.stack([out["d_real"], out["d_fake"], out["g"]]), non_blocking=True)
            torch.cuda.current_stream().synchronize()

        def e2e_drain():
            pass

    for _ in range(max(3, args.warmup // 2)):
        e2e_step()
    e2e_drain()

    def e2e_region():
        for _ in range(args.steps):
            e2e_step()
        e2e_drain()          # the last iteration's losses are read inside the timed region too

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
    e2e_losses = last[0] if graphed is not None else None

    if rank != 0:
        return
    peaks = measured_peaks()
    fam = {k: v for k, v in per_kernel.items() if "conv" in k}
    roofline = probe
    if roofline is not None and fam:
        def flops(tag):
            dims = dict(kv.split("=") for kv in tag.split() if "=" in kv)
            return 2.0 * int(dims["M"]) * int(dims["N"]) * int(dims["K"])
        n = min(args.steps, 5)
        tc = {t: v for t, v in fam.items() if t.endswith(" tc")}
        roofline["peak"] = peaks["tf_burst"]
        roofline["frac"] = roofline["achieved"] / peaks["tf_burst"]
        roofline["frac_mma_pipe"] = roofline["mma_pipe_tflops"] / peaks["tf_burst"]
        roofline["peak_source"] = peaks["source"] + " bf16 dense cuBLAS, burst (kernel timed alone)"
        roofline["tensor_core_launches_per_step"] = sum(c for c, _ in tc.values()) / n
        roofline["tensor_core_gflop_per_step"] = sum(flops(t) * c for t, (c, m) in tc.items()) / n / 1e9
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sec, cores, b_step = time_cpu_oracle(3, 1)
        cpu = {"value": b_step / sec, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "3 full config-2 iterations (B=64) after 1 warm-up; oracle port, torch %s CPU, %s; "
                         "reference pin (py2.7/PyTorch@065c5986) not installable offline" % (torch.__version__,
                                                                                           cpu_model())}
    act_mb = 4 * (13.52e6 + 12.29e6) * 2 / 1e6
    print(json.dumps({
        "metric": "G-LIS train images/sec at 80x80 bs64/GPU", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (split-bf16 tcgen05, fp32-faithful) + f32", "bf16": "bf16"}[prec_name],
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": B * world, "parallelism": "dp%d" % world,
                   "lis_depth": "all", "precision": prec_name, "cuda_graph": graphed is not None,
                   "l2": "inputs larger than L2: ~%.0f MB of activations + 36 MB of weights touched per step"
                         % act_mb,
                   "gflop_per_step": GFLOP_PER_STEP,
                   "step_tflops": GFLOP_PER_STEP * world / ms_step},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e, "host_wall_ms_per_step": wall_ms / args.steps, "last_losses": e2e_losses,
                "api": "glis_b200.trainer.HostFedStepper.submit(pinned real, z_d, z_g) -> losses"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python (no CUDA graph)")
    ap.add_argument("--kernel-table", default=None, help="write the per-kernel CUDA-event table (JSON) here")
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16x3", "bf16"])
    args = ap.parse_args()
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
