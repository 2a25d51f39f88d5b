#!/usr/bin/env python
"""Benchmark of the weak-supervision hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload pairwise|layercam]

Default workload = BASELINE.json configs[1]: AlternatingDirectionCutLoss + BoundaryLoss forward+backward on
synthetic 32x2x224x224 maps with smooth RGB affinities.  One "step" = both losses (values and d/dlogits) of one
32-image batch -- one fused launch; the metric counts pixels once per loss (2*B*H*W per step).
`--workload layercam` runs configs[2] instead (fused LayerCAM->normalise->threshold over the 3680-image set at
512x512, image-sharded, strong scaling); `--workload trainstep` config 4; the default run reports both too, at
every N, under "also".

One JSON line on stdout (rank 0).  Under torchrun every rank processes its own batch (weak scaling, no
data-path collective); the elapsed time is the max over ranks.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAIR = dict(B=32, C=2, H=224, W=224, window=5, sigma_cut=0.05, sigma_bnd=0.1, sigma_space=5.0)
LCAM = dict(S=512, layers=((1024, 32, 32), (2048, 32, 32)), chunk=128, thresh=0.3, alpha=1.0, n_images=3680)
N_SETS = 8  # rotated input sets: 8 x 45 MB > 126 MB L2
# `config` is the same object in the B200 arm and in the reference arm: both run THIS workload, batch and all
PAIR_CONFIG = {
    "workload": "configs[1]: AlternatingDirectionCutLoss + BoundaryLoss fwd/bwd, 32x2x224x224 per GPU, smooth RGB affinities "
                "(SURVEY.md 8d)",
    "batch": "32 images per step: cut loss fwd+bwd on the logits (sigma 0.05) + boundary loss fwd+bwd on softmax(logits) of "
             "every image (sigma 0.1 / 5); pixels counted once per loss (2 x 32 x 224 x 224 per step)",
    "sharding": "one batch per rank, no data-path collective",
}
PHYS_BYTES_PER_PIX = 28  # what ONE fused launch must move per pixel: 20 B in (2 logits + 3 rgb), 8 B gradient out
MIN_TIMED_MS = 20.0      # a short --steps run repeats its K steps until the timed region is at least this long
BYTES_PER_PIX = 28  # 20 B read (2 logits + 3 rgb) + 8 B gradient written, C=2
L2_MB = 126
NCU_TRAFFIC_BYTES = 32_220_000  # dram read 32.19 MB + write 0.03 MB per launch (ncu --set full, profiles/r02_c_ncu_full_pairwise.txt)
# warp instructions per launch (smsp__inst_executed.sum, profiles/r02_c_ncu_full_pairwise.txt; cut / boundary: round 1): the kernels are bound by
# instruction issue (FP32 + MUFU), not HBM -- 148 SMs x 4 schedulers x 1 instruction per clock is the second roof
NCU_WARP_INSTR = {"fused": 14.52e6, "cut": 12.13e6, "boundary": 14.11e6}
# third roof of the fused launch: MUFU (ex2 / rcp) warp instructions per launch from the same capture x 32 lanes against
# the measured MUFU rate of this B200 (tests/native/pipe_probe.cu: 4.63 T lane-op/s = 16 per clock per SM)
NCU_MUFU_WARP_INSTR = {"fused": 1.674e6}
MUFU_LANE_OPS_PER_S = 4.63e12


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def smooth_images(gen, B, H, W, device):
    import torch

    raw = torch.rand(B, 3, H + 16, W + 16, generator=gen, device=device)
    img = torch.nn.functional.avg_pool2d(raw, 9, stride=1)[..., 4:H + 4, 4:W + 4]
    lo = img.amin(dim=(1, 2, 3), keepdim=True)
    hi = img.amax(dim=(1, 2, 3), keepdim=True)
    return ((img - lo) / (hi - lo)).contiguous()


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.ok:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def _bind_to_gpu_numa_node(index):
    """Pins this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the end-to-end
    leg are first-touched on the GPU's own NUMA node (8 ranks share the host's memory controllers)."""
    if os.environ.get("WSDL_NO_NUMA_BIND") == "1":
        return
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:  # affinity is an optimisation, never a requirement
        pass


_REAL_STDOUT = None


def _emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ------------------------------------------------------------------------------------------ reference arm
def cpu_pairwise_sample(sample_B, reps, threads):
    """The reference's CPU path for configs[1] (oracle port: same ATen op sequence as
    AlternatingDirectionCutLoss.py:71-105 / AlternatingDirectionBoundaryLoss.py:20-70 incl. autograd backward)."""
    import torch

    from oracle import wsdl_oracle as O

    torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(1)
    H, W = PAIR["H"], PAIR["W"]
    logits = torch.randn(sample_B, 2, H, W, generator=gen)
    img = smooth_images(gen, sample_B, H, W, "cpu")
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        O.loss_and_grad(O.cut_loss, logits, img, sigma_color=PAIR["sigma_cut"], window_size=PAIR["window"])
        probs = torch.softmax(logits, dim=1)
        for b in range(sample_B):  # the reference's boundary loss is per image (BoundaryLoss.py:20)
            O.loss_and_grad(O.boundary_loss, probs[b], img[b], sigma_color=PAIR["sigma_bnd"],
                            sigma_space=PAIR["sigma_space"], window_size=PAIR["window"])
        best = min(best, time.perf_counter() - t0)
    return 2 * sample_B * H * W / best / 1e9, best


def cpu_layercam_sample(n_images, reps, threads):
    """The reference's CPU path for configs[2] after the backbone (oracle port of LayerCAM.py:52-76 + the threshold
    idiom of PsuedoMasks.py:59-62), one image at a time as the reference does, 512x512-sized hooks."""
    import torch

    from oracle import wsdl_oracle as O

    torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(1)
    S, layers = LCAM["S"], LCAM["layers"]
    acts = [torch.randn(n_images, C, h, w, generator=gen).relu_() for (C, h, w) in layers]
    grads = [torch.randn(n_images, C, h, w, generator=gen).mul_(1e-3) for (C, h, w) in layers]
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        for i in range(n_images):
            cam = O.layercam_from_hooks([a[i:i + 1] for a in acts], [g[i:i + 1] for g in grads], (S, S), alpha=LCAM["alpha"])
            O.threshold_mask(cam[0], LCAM["thresh"])
        best = min(best, time.perf_counter() - t0)
    return n_images / best, best


def run_reference_layercam(args):
    import torch

    threads = os.cpu_count() or 1
    n_img = 4
    times = []
    for i in range(args.warmup + args.steps):
        _, t = cpu_layercam_sample(n_img, 1, threads)
        if i >= args.warmup:
            times.append(t)
    value = n_img * len(times) / sum(times)
    _emit({
        "impl": "reference", "metric": "pseudo-masks/s", "value": value, "unit": "masks/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(times) / len(times) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[2]: fused LayerCAM->normalise->threshold over a synthetic 3680-image Pet-sized set at "
                               "512x512, image-sharded i % world == rank; one step = one pass over the set",
                   "l2": "2 rotated resident chunks of 3.2 GB per rank"},
        "how": f"{n_img} images per step, one at a time as the reference does (a bounded sample of the 3680-image pass)",
        "cpu_baseline": {"value": value, "unit": "masks/s", "cores": threads, "kind": "port",
                         "sample": f"{n_img} images per step; oracle port of LayerCAM.py:52-76 + PsuedoMasks.py:59-62 "
                                   f"(torch {torch.__version__} CPU)"},
        "e2e": {"value": value, "unit": "masks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "layercam":
        return run_reference_layercam(args)
    import torch

    threads = os.cpu_count() or 1
    sample_B = PAIR["B"]  # the whole 32-image batch of the B200 arm: same configuration, not a sample
    times = []
    for i in range(args.warmup + args.steps):
        _, t = cpu_pairwise_sample(sample_B, 1, threads)
        if i >= args.warmup:
            times.append(t)
    total = sum(times)
    pix = 2 * sample_B * PAIR["H"] * PAIR["W"] * len(times)
    value = pix / total / 1e9
    line = {
        "impl": "reference", "metric": "cut+boundary loss fwd+bwd throughput", "value": value, "unit": "Gpix/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(PAIR_CONFIG),
        "how": "the reference's own op sequence (oracle port of AlternatingDirectionCutLoss.py:71-105 and "
               "AlternatingDirectionBoundaryLoss.py:20-70 incl. autograd backward) on the host cores, the full 32-image batch "
               "per step; /root/reference does not exist on the GPU box and is a directory of scripts without an installer",
        "cpu_baseline": {"value": value, "unit": "Gpix/s", "cores": threads, "kind": "port",
                         "sample": f"{sample_B}x2x224x224 per step (the full batch), oracle port of the reference ATen op "
                                   f"sequence (torch {torch.__version__} CPU, autograd backward)"},
        "e2e": {"value": value, "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ------------------------------------------------------------------------------------------ B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pairwise", choices=["pairwise", "layercam", "pseudomask", "trainstep"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--one-stream", action="store_true")
    ap.add_argument("--separate", action="store_true", help="two launches per step (cut, boundary) instead of the fused one")
    ap.add_argument("--stream-pairs", type=int, default=2)
    ap.add_argument("--graph-steps", type=int, default=32, help="steps captured per CUDA graph (pairwise workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    args = ap.parse_args()
    if args.steps is None:  # defaults that finish within minutes for either arm
        args.steps = 30 if args.impl == "reference" else 4000
    if args.warmup is None:
        args.warmup = 3 if args.impl == "reference" else 200
    # ONE JSON line on stdout: anything a library prints there (NCCL's version banner, ...) goes to stderr instead
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from weaklysuperviseddl_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        _bind_to_gpu_numa_node(local_rank)  # before any pinned allocation: host buffers land next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _native.lib()
    stream = torch.cuda.current_stream(dev)

    if args.workload == "layercam":
        res = bench_layercam(args, lib, dev, rank, world)
    elif args.workload == "pseudomask":
        res = bench_pseudomask(args, dev, rank, world)
    elif args.workload == "trainstep":
        res = bench_trainstep(args, dev, rank, world)
    else:
        res = bench_pairwise(args, lib, dev, rank, world)
    if rank == 0:
        _emit(res)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _timed(run_steps, warmup, steps, dev, world, sampler):
    """W warm-up steps, barrier+sync, EXACTLY `steps` steps between CUDA events, sync+barrier; max over ranks."""
    import torch
    import torch.distributed as dist

    run_steps(warmup)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.start()
    e0.record()
    run_steps(steps)
    e1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, clocks


def make_pairwise_state(lib, dev, seed):
    import torch

    B, C, H, W = PAIR["B"], PAIR["C"], PAIR["H"], PAIR["W"]
    gen = torch.Generator(device=dev).manual_seed(seed)
    sets = []
    for _ in range(N_SETS):
        logits = torch.randn(B, C, H, W, generator=gen, device=dev)
        img = smooth_images(gen, B, H, W, dev)
        sets.append((logits, torch.softmax(logits, dim=1), img, torch.empty_like(logits), torch.empty_like(logits)))
    nws = lib.wsdl_pairwise_workspace_bytes(B, H, W)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    _native_check(lib.wsdl_pairwise_workspace_init(ws.data_ptr(), nws, torch.cuda.current_stream(dev).cuda_stream))
    loss_cut = torch.empty(1, device=dev)
    loss_bnd = torch.empty(B, device=dev)
    return sets, ws, nws, loss_cut, loss_bnd


def bench_pairwise(args, lib, dev, rank, world):
    import torch

    B, C, H, W = PAIR["B"], PAIR["C"], PAIR["H"], PAIR["W"]
    sets, ws, nws, loss_cut, loss_bnd = make_pairwise_state(lib, dev, 1 + rank)
    sp = lambda: torch.cuda.current_stream(dev).cuda_stream

    # The two launches of a step are independent (cut loss on the logits, boundary loss on the probabilities): they go
    # to two streams, each with its own prepared workspace, so that CTAs of one kernel fill the SM slots the other
    # leaves idle while it ramps up, waits for tiles or drains.  `--one-stream` keeps them on one stream.
    two = not args.one_stream
    n_pairs = max(1, args.stream_pairs) if two else 1  # consecutive steps rotate over this many stream pairs
    lanes = []
    for _ in range(n_pairs):
        st = (torch.cuda.Stream(dev), torch.cuda.Stream(dev)) if two else (None, None)
        wss = [torch.empty(nws, dtype=torch.uint8, device=dev) for _ in range(2)]
        for w in wss:
            _native_check(lib.wsdl_pairwise_workspace_init(w.data_ptr(), nws, torch.cuda.current_stream(dev).cuda_stream))
        lanes.append((st, wss, torch.empty(1, device=dev), torch.empty(B, device=dev)))
    torch.cuda.synchronize(dev)

    def launch_cut(i, stream_ptr, w, out=None):
        logits, probs, img, g_cut, g_bnd = sets[i % N_SETS]
        _native_check(lib.wsdl_pairwise_fwd_bwd_prepared(
            logits.data_ptr(), img.data_ptr(), B, C, H, W, PAIR["window"], PAIR["sigma_cut"], 0.0, 1, 1, 0, None,
            (out if out is not None else loss_cut).data_ptr(), g_cut.data_ptr(), w.data_ptr(), nws, stream_ptr))

    def launch_bnd(i, stream_ptr, w, out=None):
        logits, probs, img, g_cut, g_bnd = sets[i % N_SETS]
        _native_check(lib.wsdl_pairwise_fwd_bwd_prepared(
            probs.data_ptr(), img.data_ptr(), B, C, H, W, PAIR["window"], PAIR["sigma_bnd"], PAIR["sigma_space"], 0, 0, 1,
            None, (out if out is not None else loss_bnd).data_ptr(), g_bnd.data_ptr(), w.data_ptr(), nws, stream_ptr))

    fused = not args.separate  # one launch computes both losses of a step (they share the probability map)
    n_streams = 2 * n_pairs if two else 1
    fl = []  # per stream: (stream, workspace, loss_cut, loss_bnd)
    if fused:
        nwd = lib.wsdl_pairwise_dual_workspace_bytes(B, H, W)
        for k in range(n_streams):
            w = torch.empty(nwd, dtype=torch.uint8, device=dev)
            _native_check(lib.wsdl_pairwise_workspace_init(w.data_ptr(), nwd, torch.cuda.current_stream(dev).cuda_stream))
            fl.append((torch.cuda.Stream(dev) if two else None, w, torch.empty(1, device=dev), torch.empty(B, device=dev)))
        torch.cuda.synchronize(dev)

    def launch_dual(i, stream_ptr, lane):
        logits, probs, img, g_cut, g_bnd = sets[i % N_SETS]
        _, w, lc, lb = lane
        _native_check(lib.wsdl_pairwise_dual_fwd_bwd(
            logits.data_ptr(), img.data_ptr(), B, H, W, PAIR["window"], PAIR["sigma_cut"], PAIR["sigma_bnd"],
            PAIR["sigma_space"], None, None, lc.data_ptr(), lb.data_ptr(), g_cut.data_ptr(), w.data_ptr(), nwd, 1, stream_ptr))

    def some_steps(idx):  # fork from the current stream, launch, join back
        cur = torch.cuda.current_stream(dev)
        if fused:
            if not two:
                for i in idx:
                    launch_dual(i, cur.cuda_stream, fl[0])
                return
            for lane in fl:
                lane[0].wait_stream(cur)
            for i in idx:
                lane = fl[i % n_streams]
                launch_dual(i, lane[0].cuda_stream, lane)
            for lane in fl:
                cur.wait_stream(lane[0])
            return
        if not two:
            for i in idx:
                launch_cut(i, cur.cuda_stream, ws)
                launch_bnd(i, cur.cuda_stream, ws)
            return
        for (sc, sb), _, _, _ in lanes:
            sc.wait_stream(cur)
            sb.wait_stream(cur)
        for i in idx:
            (sc, sb), wss, lc, lb = lanes[i % n_pairs]
            launch_cut(i, sc.cuda_stream, wss[0], lc)
            launch_bnd(i, sb.cuda_stream, wss[1], lb)
        for (sc, sb), _, _, _ in lanes:
            cur.wait_stream(sc)
            cur.wait_stream(sb)

    graph = None
    G_STEPS = max(1, min(args.graph_steps, args.steps))  # a short run still replays one graph instead of launching directly
    if not args.no_graph:
        some_steps(range(G_STEPS))
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            some_steps(range(G_STEPS))

    def run_steps(n):
        full, rem = (n // G_STEPS, n % G_STEPS) if graph is not None else (0, n)
        for _ in range(full):
            graph.replay()
        if rem:
            some_steps(range(rem))

    # A K-step region of this workload is 0.5 ms at K = 20: too short for a stable number (and for NVML to see).  The
    # timed region is `repeats` x K steps -- the same K-step sequence over and over -- and everything below is per step.
    probe_ms, _ = _timed(run_steps, args.warmup, args.steps, dev, world, None)
    repeats = max(1, int(MIN_TIMED_MS / max(probe_ms, 1e-3)) + 1)
    sampler = ClockSampler(dev.index) if rank == 0 else None
    ms, clocks = _timed(lambda n: [run_steps(args.steps) for _ in range(n // args.steps)], 0, args.steps * repeats, dev, world,
                        sampler)
    timed_steps = args.steps * repeats
    if rank == 0 and clocks is not None and ms < 400.0:  # still short for NVML's sampling period: probe the same load, untimed
        probe = ClockSampler(dev.index)
        probe.start()
        t0 = time.time()
        while time.time() - t0 < 0.6:
            run_steps(N_SETS * 16)
            torch.cuda.synchronize(dev)
        clocks = dict(probe.stop(), note="sampled over an untimed 0.6 s repeat of the same steps (timed region < 0.4 s)")
    pix_per_step = 2 * B * H * W  # counted once per loss
    value = world * pix_per_step * timed_steps / (ms * 1e-3) / 1e9
    ms_per_step = ms / timed_steps
    peak, peak_src = peaks()
    launches_per_step = 1 if fused else 2
    launch_ms = ms_per_step / launches_per_step
    # algorithmic bytes: 28 B per pixel PER LOSS; the fused launch processes 2 B H W loss-pixels
    alg_bytes = BYTES_PER_PIX * B * H * W * (2 if fused else 1)
    phys_bytes = PHYS_BYTES_PER_PIX * B * H * W  # what a launch physically has to move (inputs once, one gradient)
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965
    instr = NCU_WARP_INSTR["fused"] if fused else 0.5 * (NCU_WARP_INSTR["cut"] + NCU_WARP_INSTR["boundary"])
    issue_floor_ms = instr / (148 * 4 * sm_mhz * 1e6) * 1e3
    per_kernel = {}
    for name, which in (("cut", 0), ("boundary", 1), ("fused cut+boundary", 2)):  # each launch alone, one stream
        if which == 2 and not fused:
            continue

        def only(n, which=which):
            for i in range(n):
                if which == 2:
                    launch_dual(i, sp(), fl[0])
                else:
                    (launch_cut if which == 0 else launch_bnd)(i, sp(), ws)
        t_ms, _ = _timed(only, 20, 500, dev, world, None)
        per_kernel[name] = t_ms / 500
    alone_ms = per_kernel["fused cut+boundary"] if fused else 0.5 * (per_kernel["cut"] + per_kernel["boundary"])
    gbs = lambda nbytes, t_ms: nbytes / (t_ms * 1e-3) / 1e9
    fractions = {
        "survey_bytes_at_step_time": gbs(alg_bytes, launch_ms) / peak,
        "survey_bytes_at_kernel_duration": gbs(alg_bytes, alone_ms) / peak,
        "physical_bytes_at_step_time": gbs(phys_bytes, launch_ms) / peak if fused else None,
        "physical_bytes_at_kernel_duration": gbs(phys_bytes, alone_ms) / peak if fused else None,
        "definitions": "survey bytes = SURVEY.md 8(d): 28 B per pixel per LOSS x the loss-pixels a launch processes (the fused "
                       "launch computes both losses: 89.9 MB); physical bytes = what that one launch must move (20 B in + 8 B "
                       "gradient out per pixel: 45.0 MB); step time = timed region / steps with consecutive steps overlapping on "
                       f"{'4 streams' if two else 'one stream'} (what a stream of independent batches sees); kernel duration = the same launch "
                       "alone on one stream, back to back, 500 launches between CUDA events (what a training step sees); all "
                       f"over the {'measured' if 'measured' in peak_src else 'fallback'} HBM peak",
    }
    res = {
        "metric": "cut+boundary loss fwd+bwd throughput", "value": value, "unit": "Gpix/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(PAIR_CONFIG),
        "how": {
            "step": ("1 fused launch (wsdl_pairwise_dual_fwd_bwd): both loss values and d/dlogits" if fused else
                     "1 cut fwd+bwd launch + 1 boundary fwd+bwd launch"),
            "timed": f"{repeats} x {args.steps} steps between one pair of CUDA events ({ms:.2f} ms); ms_per_step = region / "
                     f"{timed_steps}",
            "l2": f"rotating {N_SETS} input sets ({N_SETS * 45} MB) > {L2_MB} MB L2",
            "launch": (f"CUDA graph of {G_STEPS} steps" if graph is not None else "direct C-ABI calls") +
                      ((f"; consecutive steps rotate over {n_streams} streams (independent work, own workspaces)" if fused else
                        f"; the cut and the boundary launch of a step on two streams, consecutive steps on {n_pairs} stream "
                        "pairs (independent work, own workspaces)") if two else "; one stream"),
        },
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_TRAFFIC_BYTES,
                     "kernel": "pairwise_dual_kernel" if fused else "pairwise_sym_kernel<2,true|false>",
                     "algorithmic_bytes_per_launch": alg_bytes, "physical_bytes_per_launch": phys_bytes if fused else alg_bytes,
                     "launch_ms": launch_ms, "kernel_ms_alone": alone_ms, "fractions": fractions,
                     "peak_source": peak_src, "per_kernel_ms_direct_launch": per_kernel,
                     "issue_roof": {"warp_instructions_per_launch": instr, "floor_ms": issue_floor_ms,
                                    "frac": issue_floor_ms / launch_ms,
                                    "note": "second roof: ncu warp instructions / (148 SMs x 4 schedulers x SM clock)"},
                     **({"mufu_roof": {"warp_instructions_per_launch": NCU_MUFU_WARP_INSTR["fused"],
                                       "floor_ms": NCU_MUFU_WARP_INSTR["fused"] * 32 / MUFU_LANE_OPS_PER_S * 1e3,
                                       "frac": NCU_MUFU_WARP_INSTR["fused"] * 32 / MUFU_LANE_OPS_PER_S * 1e3 / launch_ms,
                                       "note": "third roof: MUFU.EX2 + MUFU.RCP lane-ops / measured 4.63 T/s; the march "
                                               "needs 768 MUFU cycles, ~670 issue slots and ~700 FP32-pipe cycles per "
                                               "warp step: the three pipes are balanced"}} if fused else {}),
                     "note": "`frac` = survey bytes at step time (the contract's definition: algorithmic bytes / average launch "
                             "duration over the timed region; launches of consecutive steps overlap on the device).  The same "
                             "bytes at the kernel's own duration and the physical bytes at both are in `fractions`.  traffic = "
                             "dram__bytes_read+write per launch (ncu --set full, profiles/): below the algorithmic bytes because the "
                             "12.8 MB gradient is still dirty in L2 when the kernel ends; the kernel is bound by MUFU / FP32 / "
                             "issue, not HBM (DESIGN.md 4.2)"},
        "gpu_launches": launches_per_step * timed_steps,
        "clocks": clocks,
    }
    e2e = e2e_pairwise(dev, world, max(3, min(50, args.steps)), max(3, min(5, args.warmup)))
    if rank == 0:
        res["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, t = cpu_pairwise_sample(PAIR["B"], 2, threads)
            res["cpu_baseline"] = {"value": v, "unit": "Gpix/s", "cores": threads, "kind": "port",
                                   "sample": f"the full 32x2x224x224 batch, best of 2, {t:.2f} s per step; oracle port of the "
                                             "reference ATen op sequence incl. autograd backward"}
    if not args.no_also:  # configs 3 and 4 ride along at every N (collective: every rank runs them, rank 0 reports)
        also = {}
        for key, fn in (("layercam_3680", lambda: bench_layercam_sharded(dev, rank, world, 3, 1)),
                        ("trainstep", lambda: _trainstep_brief(bench_trainstep(argparse.Namespace(
                            steps=16, warmup=8, no_cpu_baseline=True), dev, rank, world)))):
            try:
                also[key] = fn()
            except Exception as e:  # a side measurement must never take the headline line down
                also[key] = {"error": repr(e)[:300]}
        if world == 1:
            try:
                also["eager_torch_cuda"] = eager_torch_pairwise(dev)
            except Exception as e:
                also["eager_torch_cuda"] = {"error": repr(e)[:300]}
        if rank == 0:
            res["also"] = also
    return res


def _trainstep_brief(r):
    return {k: r[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "ms_per_step", "scaling", "stage_share") if k in r} | {
        "config": r["config"]["workload"], "e2e_h2d_bytes_per_step": r["e2e"]["h2d_bytes_per_step"]}


def eager_torch_pairwise(dev):
    """SURVEY 8d's second bar: the reference's math as plain eager PyTorch ON THE SAME GPU (the way the reference
    itself would run there: 24 shifted copies per loss, autograd backward, the boundary loss image by image), timed on a
    quarter of the configs[1] batch and cross-checked against the fused launch.  Written out here (not imported from
    oracle/) -- a timing comparator, not a parity check."""
    import torch
    import torch.nn.functional as F

    from weaklysuperviseddl_b200 import functional as WF

    def pair_loss(p, img, sigma_color, sigma_space, per_channel_mean):
        # p (B,C,H,W) probabilities, img (B,3,H,W); AlternatingDirectionCutLoss.py:71-105 / ...BoundaryLoss.py:20-70
        pad = 2
        H, W = p.shape[-2:]
        pp, ip = F.pad(p, (pad,) * 4, mode="reflect"), F.pad(img, (pad,) * 4, mode="reflect")
        total, K = 0.0, 0
        for dy in range(-pad, pad + 1):
            for dx in range(-pad, pad + 1):
                if dy == 0 and dx == 0:
                    continue
                ps = pp[:, :, pad + dy:pad + dy + H, pad + dx:pad + dx + W]
                isf = ip[:, :, pad + dy:pad + dy + H, pad + dx:pad + dx + W]
                e = -((img - isf) ** 2).sum(1, keepdim=True) / (2 * sigma_color ** 2)
                if sigma_space:
                    e = e - (dx * dx + dy * dy) / (2 * sigma_space ** 2)
                d2 = (p - ps) ** 2
                total = total + (torch.exp(e) * (d2.mean(1, keepdim=True) if per_channel_mean else d2.sum(1, keepdim=True))).mean()
                K += 1
        return total / K

    Bq = PAIR["B"] // 4
    g = torch.Generator(device=dev).manual_seed(77)
    logits = torch.randn(Bq, 2, PAIR["H"], PAIR["W"], device=dev, generator=g)
    img = smooth_images(g, Bq, PAIR["H"], PAIR["W"], dev)

    def step():
        x = logits.clone().requires_grad_(True)
        cut = pair_loss(torch.softmax(x, 1), img, PAIR["sigma_cut"], None, True)
        probs = torch.softmax(x, 1)
        bnd = torch.stack([pair_loss(probs[b:b + 1], img[b:b + 1], PAIR["sigma_bnd"], PAIR["sigma_space"], False)
                           for b in range(Bq)])
        (cut + bnd.sum()).backward()
        return cut.detach(), bnd.detach(), x.grad

    cut, bnd, grad = step()
    lc, lb, gk = WF.pairwise_dual_loss_and_grad(logits, img, PAIR["sigma_cut"], PAIR["sigma_bnd"], PAIR["sigma_space"], 5)
    agree = {"cut_rel": abs(cut.item() - lc.item()) / abs(cut.item()),
             "boundary_rel": ((bnd - lb).abs().max() / bnd.abs().max()).item(),
             "grad_rel": ((grad - gk).abs().max() / grad.abs().max()).item()}
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    return {"metric": "cut+boundary loss fwd+bwd, eager PyTorch on the same B200", "value": 2 * Bq * PAIR["H"] * PAIR["W"] / (ms * 1e-3) / 1e9,
            "unit": "Gpix/s", "ms_per_step": ms, "sample": f"{Bq} of {PAIR['B']} images per step, {reps} steps",
            "agreement_with_fused_launch": agree}


def _native_check(rc):
    if rc != 0:
        from weaklysuperviseddl_b200 import _native

        _native.check(rc, "wsdl call in bench.py")


def e2e_pairwise(dev, world, steps, warmup):
    """Same metric through the public API with HOST buffers: every step copies its logits and images from pinned host
    memory, runs the loss forward + backward, and copies the gradient and the loss values back.  Three variants:

      u8   (the headline): images as the dataset stores them, 8-bit RGB; the kernel reads value / 255 exactly as
           ToTensor does, so the result is bit-identical to the f32-image call on the same pixels -- 3 B instead of 12 B
           per pixel over PCIe.  f32 logits in, f32 gradient out.  One launch: functional.weak_loss_and_grad
           (wsdl_weak_loss_fwd_bwd).
      f32  round 1's path: f32 images, LocalNormalizedCutLoss + ConstrainToBoundaryLossSingle modules + .backward().
      bf16 u8 images, bf16 logits in, bf16 gradient out (what a bf16-autocast network hands over and takes back).

    Every rank runs its own batch concurrently (shared host links included); max time over ranks.  (Capturing K of these
    pipelined steps in a CUDA graph was measured too: 6.1 vs 7.4 Gpix/s -- the step is bound by the PCIe link carrying
    41 GB/s in and 30 GB/s out at once, not by the host's launch overhead.)"""
    import torch
    import torch.distributed as dist

    import weaklysuperviseddl_b200 as Wm
    from weaklysuperviseddl_b200 import functional as WF

    B, C, H, W = PAIR["B"], PAIR["C"], PAIR["H"], PAIR["W"]
    gen = torch.Generator().manual_seed(1)
    logits32 = torch.randn(B, C, H, W, generator=gen)
    img8 = (smooth_images(gen, B, H, W, "cpu") * 255).round().to(torch.uint8)
    cut = Wm.LocalNormalizedCutLoss(PAIR["sigma_cut"], PAIR["window"])
    bnd = Wm.ConstrainToBoundaryLossSingle(PAIR["sigma_bnd"], PAIR["sigma_space"], PAIR["window"])
    main = torch.cuda.current_stream(dev)
    go_c = torch.ones(1, device=dev)
    go_b = torch.full((B,), 1.0 / B, device=dev)

    def run(kind):
        h_logits = (logits32.to(torch.bfloat16) if kind == "bf16" else logits32).pin_memory()
        h_img = (img8.float() / 255 if kind == "f32" else img8).pin_memory()
        # Three streams, three buffer sets: the H2D copy of step k+1 and the D2H read-back of step k-1 run under the
        # compute of step k (PCIe is full duplex; a third set keeps both copy queues fed: 12.5 vs 9.7 Gpix/s on the bf16
        # leg).  Every step still moves its own inputs and results.
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        NB = 3
        d_logits = [torch.empty_like(h_logits, device=dev) for _ in range(NB)]
        d_img = [torch.empty_like(h_img, device=dev) for _ in range(NB)]
        h_grads = [torch.empty_like(h_logits).pin_memory() for _ in range(NB)]
        h_losses = [torch.empty(3 + B).pin_memory() for _ in range(NB)]
        ev_in = [torch.cuda.Event() for _ in range(NB)]
        ev_done = [torch.cuda.Event() for _ in range(NB)]
        ev_out = [torch.cuda.Event() for _ in range(NB)]
        state = {"k": 0}

        def step():
            k = state["k"]
            i = k % NB
            state["k"] = k + 1
            if k >= NB:
                ev_out[i].synchronize()  # the host has the results of the step that last used this buffer set
            with torch.cuda.stream(s_in):
                if k >= NB:
                    s_in.wait_event(ev_done[i])  # its inputs are no longer being read
                d_logits[i].copy_(h_logits, non_blocking=True)
                d_img[i].copy_(h_img, non_blocking=True)
                ev_in[i].record(s_in)
            main.wait_event(ev_in[i])
            if kind == "f32":
                logits = d_logits[i].detach().requires_grad_(True)
                loss = cut(logits, d_img[i]) + bnd(torch.softmax(logits, dim=1), d_img[i]).mean()
                loss.backward()
                g, l = logits.grad, loss.detach().reshape(1).expand(3 + B).contiguous()
            else:
                l = torch.empty(3 + B, device=dev)  # total, (ce), cut, bnd[B]: one read-back
                g = WF.weak_loss_and_grad(d_logits[i], d_img[i], None, 0.0, go_c, go_b, PAIR["sigma_cut"], PAIR["sigma_bnd"],
                                          PAIR["sigma_space"], PAIR["window"], packed_out=l)[4]
            ev_done[i].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[i])
                h_grads[i].copy_(g, non_blocking=True)
                h_losses[i].copy_(l, non_blocking=True)
                g.record_stream(s_out)
                l.record_stream(s_out)
                ev_out[i].record(s_out)

        def drain():
            for e in ev_out:
                e.synchronize()
            torch.cuda.synchronize(dev)

        for _ in range(warmup):
            step()
        drain()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for _ in range(steps):
            step()
        main.wait_stream(s_out)  # the last read-backs are inside the timed region
        main.wait_stream(s_in)
        e1.record(main)
        drain()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return {"value": world * 2 * B * H * W * steps / (ms * 1e-3) / 1e9, "unit": "Gpix/s",
                "h2d_bytes_per_step": h_logits.numel() * h_logits.element_size() + h_img.numel() * h_img.element_size(),
                "d2h_bytes_per_step": h_grads[0].numel() * h_grads[0].element_size() + h_losses[0].numel() * 4,
                "steps": steps, "ms_per_step": ms / steps}, h_losses[(steps + warmup - 1) % NB].clone()

    u8, l_u8 = run("u8")
    f32, l_f32 = run("f32")
    bf16, _ = run("bf16")
    u8["api"] = ("weaklysuperviseddl_b200.functional.weak_loss_and_grad(logits f32, images u8) -> wsdl_weak_loss_fwd_bwd: cut + "
                 "mean boundary loss values and d/dlogits in one launch; pinned host buffers; H2D / compute / D2H of consecutive "
                 "steps pipelined on three streams")
    u8["images"] = "uint8 RGB as the dataset stores them (read as value / 255, bit-identical to ToTensor's floats)"
    u8["total_loss_vs_f32_image_path_rel"] = abs(float(l_u8[0]) - float(l_f32[0])) / abs(float(l_f32[0]))
    u8["variants"] = {
        "f32_images_modules": dict(f32, api="LocalNormalizedCutLoss()(logits, images) + ConstrainToBoundaryLossSingle()(softmax, "
                                           "images).mean(); .backward() -- round 1's e2e path, f32 images"),
        "bf16_logits_u8_images": dict(bf16, api="weak_loss_and_grad(logits bf16, images u8): bf16 gradient back"),
    }
    return u8


# ------------------------------------------------------------------------------------------ configs[2]
def bench_layercam_core(lib, dev, rank, world, steps, warmup):
    """Fused LayerCAM -> normalise -> upsample -> threshold on a resident chunk of 512x512-sized hooks
    (1024x32x32 + 2048x32x32 fp32 per image); two rotated chunks (2 x 3.2 GB) defeat the L2."""
    import torch

    chunk = LCAM["chunk"]
    S = LCAM["S"]
    layers = LCAM["layers"]
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    sets = []
    for _ in range(2):
        acts = [torch.randn(chunk, C, h, w, device=dev, generator=gen).relu_() for (C, h, w) in layers]
        grads = [torch.randn(chunk, C, h, w, device=dev, generator=gen).mul_(1e-3) for (C, h, w) in layers]
        sets.append((acts, grads))
    n = len(layers)
    IntArr, PtrArr = ctypes.c_int * n, ctypes.c_void_p * n
    Cs, hs, ws_ = IntArr(*[l[0] for l in layers]), IntArr(*[l[1] for l in layers]), IntArr(*[l[2] for l in layers])
    nws = lib.wsdl_layercam_workspace_bytes(Cs, hs, ws_, n, chunk, 0)
    # Two streams, each with its own workspace and mask buffer, chunks alternating: the (instruction-bound,
    # L2-resident) upsample/threshold kernel of one chunk runs under the (HBM-bound) channel sum of the next.
    main = torch.cuda.current_stream(dev)
    lanes = [(torch.cuda.Stream(dev), torch.empty(nws, dtype=torch.uint8, device=dev),
              torch.empty(chunk, S, S, dtype=torch.uint8, device=dev)) for _ in range(2)]
    near = torch.zeros(2, dtype=torch.int64, device=dev)

    def one(i):
        acts, grads = sets[i % 2]
        st, wsb, mask = lanes[i % 2]
        rc = lib.wsdl_layercam_fused(PtrArr(*[t.data_ptr() for t in acts]), PtrArr(*[t.data_ptr() for t in grads]),
                                     Cs, hs, ws_, n, chunk, 0, S, S, LCAM["alpha"], 0, LCAM["thresh"], 1e-6, None,
                                     mask.data_ptr(), near[i % 2:].data_ptr(), wsb.data_ptr(), nws, st.cuda_stream)
        _native_check(rc)

    def run_steps(k):  # called with the events of _timed recorded on `main`: fork to the two lanes, join back
        for st, _, _ in lanes:
            st.wait_stream(main)
        for i in range(k):
            one(i)
        for st, _, _ in lanes:
            main.wait_stream(st)

    ms, _ = _timed(run_steps, warmup, steps, dev, world, None)
    per_image = sum(2 * C * h * w * 4 for (C, h, w) in layers) + S * S
    masks_per_s = world * chunk * steps / (ms * 1e-3)
    peak, peak_src = peaks()
    achieved = per_image * chunk * steps / (ms * 1e-3) / 1e9
    return {"metric": "pseudo-masks/s (fused LayerCAM->normalise->threshold, 512x512, layers 3+4 fp32)",
            "value": masks_per_s, "unit": "masks/s", "gpix_per_s": masks_per_s * S * S / 1e9,
            "ms_per_step": ms / steps, "images_per_step": chunk, "steps": steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "algorithmic_bytes_per_image": per_image, "peak_source": peak_src,
                         "note": "both kernels of the call (channel-sum + upsample/threshold) and the ctrl memset; "
                                 "consecutive chunks alternate between two streams; peak = copy-measured HBM bandwidth "
                                 "(a read-only stream can exceed it)"},
            "time_for_3680_images_ms": LCAM["n_images"] / (masks_per_s / world) * 1e3 / world,
            "l2": "2 rotated chunks of 3.2 GB", "streams": 2}


def bench_layercam_sharded(dev, rank, world, passes, warmup_passes):
    """BASELINE config 3 as north_star states it: the fused LayerCAM -> normalise -> threshold kernels over the WHOLE
    synthetic 3680-image Pet-sized set at 512x512 (layer3 + layer4 hooks, fp32), image-sharded i % world == rank through
    the product entry `generate_pseudo_masks_sharded` (resident chunks of 128 images, two streams, counters all-reduced
    at the end).  STRONG scaling: the set is fixed, every rank walks its 3680 / world images.  The hooks of a chunk are
    synthetic and resident (two rotated 3.2 GB chunks defeat the L2; generating 93.6 GB of random numbers is not part of
    the path); one pass = the whole set."""
    import torch

    from weaklysuperviseddl_b200.PsuedoMasks import generate_pseudo_masks_sharded

    chunk, S, layers, n_images = LCAM["chunk"], LCAM["S"], LCAM["layers"], LCAM["n_images"]
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    sets = []
    for _ in range(2):
        acts = [torch.randn(chunk, C, h, w, device=dev, generator=gen).relu_() for (C, h, w) in layers]
        grads = [torch.randn(chunk, C, h, w, device=dev, generator=gen).mul_(1e-3) for (C, h, w) in layers]
        sets.append((acts, grads))
    calls = {"n": 0}

    def hooks(indices):
        acts, grads = sets[calls["n"] % 2]
        calls["n"] += 1
        n = len(indices)
        return [a[:n] for a in acts], [g[:n] for g in grads]

    last = {}

    def one_pass():
        last["res"] = generate_pseudo_masks_sharded(hooks, n_images, (S, S), cam_thresh=LCAM["thresh"], alpha=LCAM["alpha"],
                                                    chunk=chunk, rank=rank, world=world, streams=2, count_foreground=False,
                                                    sink=lambda idx, m: last.__setitem__("mask", m))

    ms, _ = _timed(lambda n: [one_pass() for _ in range(n)], warmup_passes, passes, dev, world, None)
    per_image = sum(2 * C * h * w * 4 for (C, h, w) in layers) + S * S
    masks_per_s = n_images * passes / (ms * 1e-3)
    peak, peak_src = peaks()
    achieved = per_image * masks_per_s / 1e9 / world  # per GPU
    return {"metric": "pseudo-masks/s over the 3680-image set (fused LayerCAM->normalise->threshold, 512x512, layers 3+4 fp32)",
            "value": masks_per_s, "unit": "masks/s", "n_gpus": world, "scaling": "strong", "passes": passes,
            "ms_per_pass": ms / passes, "images": n_images, "images_per_rank": len(range(rank, n_images, world)),
            "gpix_per_s": masks_per_s * S * S / 1e9, "counters": last["res"]["counters"],
            "api": "weaklysuperviseddl_b200.generate_pseudo_masks_sharded (i % world == rank, chunks of 128, 2 streams)",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "algorithmic_bytes_per_image": per_image, "peak_source": peak_src,
                         "note": "per GPU: algorithmic bytes of the images a rank processed / the pass time (max over ranks); "
                                 "both kernels of every call, the per-call control memset, the per-chunk counter sums and the "
                                 "one host read at the end of the pass"}}


def bench_layercam(args, lib, dev, rank, world):
    passes = max(1, min(args.steps, 10))
    core = bench_layercam_sharded(dev, rank, world, passes, 1)
    chunk_core = bench_layercam_core(lib, dev, rank, world, 30, 5)  # round 1's resident-chunk loop, for continuity
    # end to end for this stage = images from the host through the classifier's hooks (cuDNN) to masks on the host
    pm = bench_pseudomask(argparse.Namespace(steps=20, warmup=4, no_cpu_baseline=True), dev, rank, world, S=LCAM["S"], B=16,
                          cpu_baseline=False)
    return {
        "metric": "pseudo-masks/s", "value": core["value"], "unit": "masks/s", "n_gpus": world, "steps": passes,
        "warmup": 1, "ms_per_step": core["ms_per_pass"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[2]: fused LayerCAM->normalise->threshold over a synthetic 3680-image Pet-sized set at "
                               "512x512, image-sharded i % world == rank; one step = one pass over the set",
                   "l2": "2 rotated resident chunks of 3.2 GB per rank"},
        "roofline": core["roofline"], "gpu_launches": 2 * passes * ((core["images_per_rank"] + LCAM["chunk"] - 1) // LCAM["chunk"]),
        "extra": {"sharded_pass": core, "resident_chunk_loop": chunk_core},
        "e2e": dict(pm["e2e"], stage_share=pm["stage_share"],
                    note="16x3x512x512 host images per step -> H2D -> ResNet-50 forward + backward to the layer3/layer4 hooks "
                         "(cuDNN, out of scope) -> fused LayerCAM -> mask -> keep_largest -> D2H masks"),
        **({"cpu_baseline": _layercam_cpu_baseline()} if (world == 1 and rank == 0 and not args.no_cpu_baseline) else {}),
    }


def _layercam_cpu_baseline():
    threads = os.cpu_count() or 1
    v, t = cpu_layercam_sample(4, 2, threads)
    return {"value": v, "unit": "masks/s", "cores": threads, "kind": "port",
            "sample": f"4 images of 512x512-sized hooks, one at a time, best of 2, {t:.2f} s; oracle port of "
                      "LayerCAM.py:52-76 + PsuedoMasks.py:59-62"}


# ------------------------------------------------------------------------------------------ configs[0]
def _hookable_resnet50(num_classes=37):
    """The model contract of the reference's classifier (ClassificationModel.py:9-41): ResNet-50 with a dilated last
    stage, frozen weights, hookable layer2/3/4, forward -> (logits, [f2, f3, f4]).  Random initialisation (no network)."""
    import torch
    import torch.nn as nn
    import torchvision

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            r = torchvision.models.resnet50(weights=None, replace_stride_with_dilation=[False, False, True])
            for p in r.parameters():
                p.requires_grad = False
            self.layer0 = nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool)
            self.layer1, self.layer2, self.layer3, self.layer4 = r.layer1, r.layer2, r.layer3, r.layer4
            self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
            self.fc = nn.Linear(r.fc.in_features, num_classes)

        def forward(self, x):
            x = self.layer0(x)
            f1 = self.layer1(x)
            f2 = self.layer2(f1)
            f3 = self.layer3(f2)
            f4 = self.layer4(f3)
            return self.fc(self.avgpool(f4).flatten(1)), [f2, f3, f4]

    torch.manual_seed(0)
    return Net().eval()


def cpu_pseudomask_sample(n_images, threads):
    """The reference's path for configs[0] on the host: one image at a time through the classifier (forward + backward
    to the hooks, LayerCAM.py:35-48), LayerCAM.py:52-76, threshold (PsuedoMasks.py:59-62), keep_largest (:15-21)."""
    import torch

    from oracle import wsdl_oracle as O

    torch.set_num_threads(threads)
    net = _hookable_resnet50()
    store = {}
    for name in ("layer3", "layer4"):
        layer = getattr(net, name)
        layer.register_forward_hook(lambda m, i, o, name=name: store.__setitem__("a_" + name, o))
        layer.register_full_backward_hook(lambda m, gi, go, name=name: store.__setitem__("g_" + name, go[0]))
    gen = torch.Generator().manual_seed(0)
    imgs = torch.rand(n_images, 3, 224, 224, generator=gen)
    labels = torch.randint(0, 37, (n_images,), generator=gen)
    t0 = time.perf_counter()
    for i in range(n_images):
        x = imgs[i:i + 1].clone().requires_grad_()
        logits, _ = net(x)
        logits.gather(1, labels[i:i + 1].view(-1, 1)).squeeze().backward()
        cam = O.layercam_from_hooks([store["a_layer3"], store["a_layer4"]], [store["g_layer3"], store["g_layer4"]], (224, 224))
        O.keep_largest(O.threshold_mask(cam[0], 0.3))
    dt = time.perf_counter() - t0
    return n_images / dt, dt


def bench_pseudomask(args, dev, rank, world, S=224, B=32, cpu_baseline=True):
    """configs[0] through the drop-in classes, end to end: pinned host images -> H2D -> classifier forward + backward
    (cuDNN) -> fused LayerCAM/normalise/upsample/threshold -> keep_largest -> D2H masks.  B = 32 images per step.
    (S = 512 is the end-to-end leg of the configs[2] line.)"""
    import torch

    from weaklysuperviseddl_b200 import functional as WF
    from weaklysuperviseddl_b200.LayerCAM import LayerCAMGenerator

    torch.backends.cudnn.benchmark = True  # fixed shapes: the classifier is cuDNN's (out of scope)
    net = _hookable_resnet50().to(dev)
    gen_cam = LayerCAMGenerator(net, ["layer3", "layer4"])
    g = torch.Generator().manual_seed(1 + rank)
    h_imgs = torch.rand(B, 3, S, S, generator=g).pin_memory()
    h_masks = torch.empty(B, S, S, dtype=torch.uint8).pin_memory()
    labels = torch.randint(0, 37, (B,), generator=g).to(dev)
    t_fused = []

    def step(measure=False):
        imgs = h_imgs.to(dev, non_blocking=True)
        acts, grads = gen_cam._forward_backward(imgs, labels)
        if measure:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        cam, mask, near = WF.layercam_fused(acts, grads, (S, S), thresh=0.3, want_cam=False)
        mask = WF.keep_largest(mask)
        if measure:
            e1.record()
        h_masks.copy_(mask, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        if measure:
            t_fused.append(e0.elapsed_time(e1))

    def run_steps(n):
        for _ in range(n):
            step()

    steps = min(args.steps, 200)
    ms, _ = _timed(run_steps, max(3, min(args.warmup, 10)), steps, dev, world, None)
    for _ in range(5):
        step(measure=True)
    value = world * B * steps / (ms * 1e-3)
    res = {
        "metric": "pseudo-masks/s (end to end: classifier fwd+bwd + LayerCAM -> mask + keep_largest)", "value": value,
        "unit": "masks/s", "n_gpus": world, "steps": steps, "warmup": max(3, min(args.warmup, 10)), "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[0]: LayerCAM pseudo-mask generation, ResNet-50 (dilated layer4, random init), "
                               f"37 classes, {B}x3x{S}x{S} per step per GPU, through LayerCAMGenerator + keep_largest"},
        "e2e": {"value": value, "unit": "masks/s", "h2d_bytes_per_step": h_imgs.numel() * 4,
                "d2h_bytes_per_step": h_masks.numel(), "steps": steps, "ms_per_step": ms / steps,
                "api": "LayerCAMGenerator(model, ['layer3','layer4']) hooks + fused layercam + keep_largest; pinned host buffers"},
        "stage_share": {"fused_layercam_plus_keep_largest_ms": sum(t_fused) / len(t_fused), "step_ms": ms / steps,
                        "note": "the rest of the step is the cuDNN classifier (out of scope by north_star) and PCIe"},
        "gpu_launches": 4 * steps,
    }
    if cpu_baseline and world == 1 and rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, t = cpu_pseudomask_sample(2, threads)
        res["cpu_baseline"] = {"value": v, "unit": "masks/s", "cores": threads, "kind": "port",
                               "sample": f"2 images, one at a time, {t:.2f} s: torchvision ResNet-50 on the host cores + "
                                         "oracle port of LayerCAM.py:52-76, PsuedoMasks.py:59-62, :15-21"}
    return res


def cpu_trainstep_sample(threads):
    """One image through the same step on the host cores: DeepLabV3-R50 fp32 forward + backward and the oracle port of
    the reference's loss composition."""
    import torch
    import torch.nn.functional as F
    import torchvision

    sys.path.insert(0, ROOT)
    from oracle import wsdl_oracle as O

    torch.set_num_threads(threads)
    net = torchvision.models.segmentation.deeplabv3_resnet50(weights=None, weights_backbone=None, num_classes=2)
    net.train()
    for m in net.modules():  # batch of one: batch-norm statistics need more than one value per channel
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
    gen = torch.Generator().manual_seed(0)
    img = smooth_images(gen, 1, 224, 224, "cpu")
    labels = (img.mean(1) > img.mean()).long()
    t0 = time.perf_counter()
    out = net(img)["out"]
    p = torch.softmax(out, 1)
    loss = (F.cross_entropy(out, labels) + 0.1 * O.cut_loss(out, img, sigma_color=0.05, window_size=5)
            + 0.5 * O.boundary_loss(p[0], img[0], sigma_color=0.1, sigma_space=5, window_size=5))
    loss.backward()
    dt = time.perf_counter() - t0
    return 1.0 / dt, dt


def bench_trainstep(args, dev, rank, world):
    """BASELINE config 4: the weakly-supervised segmentation train step.  DeepLabV3-R50 (2 classes, random init, bf16
    autocast, cuDNN: out of scope by north_star) on 32x3x224x224 per GPU; pseudo-labels made ONCE (untimed) by the fused
    LayerCAM -> mask kernel from synthetic hooks; loss = CE + 0.1 cut + 0.5 mean boundary through WeakSupervisionLoss
    (ONE launch: cross-entropy + cut + boundary, wsdl_weak_loss_fwd_bwd); Adam; DistributedDataParallel over NCCL when world > 1.  Every step copies its images
    and labels from pinned host memory and reads the loss back."""
    import torch
    import torch.distributed as dist
    import torchvision

    from weaklysuperviseddl_b200 import WeakSupervisionLoss
    from weaklysuperviseddl_b200 import functional as WF

    B, S = 32, 224
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True  # fixed shapes: let cuDNN pick its kernels once (warm-up steps)
    fmt = torch.contiguous_format if os.environ.get("WSDL_TRAIN_NCHW") == "1" else torch.channels_last
    net = torchvision.models.segmentation.deeplabv3_resnet50(weights=None, weights_backbone=None, num_classes=2)
    net = net.to(dev).to(memory_format=fmt).train()
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[dev.index], gradient_as_bucket_view=True)
        if os.environ.get("WSDL_DDP_FP32") != "1":  # bf16 backbone: all-reduce the gradient buckets in bf16 as well
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks

            net.register_comm_hook(None, default_hooks.bf16_compress_hook)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    crit = WeakSupervisionLoss()
    g = torch.Generator().manual_seed(11 + rank)
    h_imgs = smooth_images(g, B, S, S, "cpu").pin_memory()
    # pseudo-labels: synthetic layer3/layer4 hooks -> fused LayerCAM -> threshold -> keep_largest (the config-3 kernels)
    gd = torch.Generator(device=dev).manual_seed(5 + rank)
    acts = [torch.relu(torch.randn(B, c, 14, 14, device=dev, generator=gd)) for c in (1024, 2048)]
    grads = [torch.randn(B, c, 14, 14, device=dev, generator=gd) * 1e-3 for c in (1024, 2048)]
    _, mask, _ = WF.layercam_fused(acts, grads, (S, S), thresh=0.3, want_cam=False)
    h_labels = WF.keep_largest(mask).cpu().pin_memory()  # uint8 {0,1}, the reference's on-disk mask after clamp(max=1)
    del acts, grads
    h_loss = torch.empty((), dtype=torch.float32).pin_memory()
    t_loss = []

    def step(measure=False):
        imgs = h_imgs.to(dev, non_blocking=True)  # NCHW for the pairwise kernels, channels-last copy for cuDNN
        labels = h_labels.to(dev, non_blocking=True)  # u8 {0,1}: the fused loss reads them as they are
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(imgs.contiguous(memory_format=fmt))["out"]
        if measure:
            out = out.detach().requires_grad_(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        total, _ = crit(out, imgs, labels)
        total.backward()
        if measure:
            e1.record()
            torch.cuda.current_stream(dev).synchronize()
            t_loss.append(e0.elapsed_time(e1))
            return
        opt.step()
        h_loss.copy_(total.detach(), non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    def run_steps(n):
        for _ in range(n):
            step()

    steps, warmup = min(args.steps, 60), max(3, min(args.warmup, 8))
    sampler = ClockSampler(dev.index) if rank == 0 else None
    ms, clocks = _timed(run_steps, warmup, steps, dev, world, sampler)
    for _ in range(6):
        step(measure=True)
    loss_ms = sorted(t_loss)[len(t_loss) // 2]
    value = world * B * steps / (ms * 1e-3)
    peak, peak_src = peaks()
    res = {
        "metric": "train-step images/s (weakly-supervised segmentation step: backbone + CE + cut + boundary + Adam)",
        "value": value, "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 backbone / f32 losses",
        "data": "synthetic",
        "config": {"workload": "config 4: DeepLabV3-R50 (2 classes, random init) bf16 autocast, 32x3x224x224 per GPU, "
                               "pseudo-labels from the fused LayerCAM->mask kernels, loss = CE + 0.1 cut + 0.5 mean boundary "
                               "(WeakSupervisionLoss: one launch for all three terms on the bf16 logits), Adam" +
                               (", DistributedDataParallel over NCCL" if world > 1 else ""),
                   "l2": "activations of the backbone (> 126 MB) pass through L2 between the loss launches"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": h_imgs.numel() * 4 + h_labels.numel(),
                "d2h_bytes_per_step": 4, "steps": steps, "ms_per_step": ms / steps,
                "api": "torchvision deeplabv3_resnet50 + weaklysuperviseddl_b200.WeakSupervisionLoss; pinned host buffers"},
        "stage_share": {"loss_fwd_bwd_ms": loss_ms, "step_ms": ms / steps, "share": loss_ms / (ms / steps),
                        "note": "CE + fused cut/boundary launch + their autograd glue, CUDA events; the rest is the cuDNN "
                                "backbone (out of scope by north_star), Adam, and for N > 1 the NCCL gradient all-reduce"},
        "roofline": {"bound": "hbm", "achieved": BYTES_PER_PIX * 2 * B * S * S / (loss_ms * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": BYTES_PER_PIX * 2 * B * S * S / (loss_ms * 1e-3) / 1e9 / peak, "traffic": None,
                     "kernel": "pairwise_dual_kernel inside the loss (one launch per step, not overlapped with another)",
                     "peak_source": peak_src},
        "gpu_launches": steps,
        "clocks": clocks,
    }
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, t = cpu_trainstep_sample(threads)
        res["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
                               "sample": f"1 image, {t:.2f} s: torchvision DeepLabV3-R50 fp32 fwd+bwd on the host cores + "
                                         "oracle port of the loss composition"}
    return res


if __name__ == "__main__":
    main()
