#!/usr/bin/env python
"""OSVOS hot-path benchmark (contract: see the task statement / DESIGN.md section 7).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload infer480|train480]

One "step" = one pass of the hot path over one batch of synthetic input:
  infer480 (default, BASELINE.json configs[1]): forward of one 480x854 frame, batch 1;
  train480 (configs[2] shape): forward + online loss + backward of one 480x854 frame.
N > 1 runs one replica per GPU over independent frames (no data-path collective for
inference / online fine-tuning; the parent-training allreduce is benchmarked separately) -> weak scaling.

Prints ONE JSON line on rank 0 with metric/value/unit, e2e (host buffers, H2D + D2H inside the
timed region, through the public nn.Module API), roofline (dominant kernel = the tcgen05 conv),
cpu_baseline (oracle port on the host cores), clocks and gpu_launches.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# Only the result line may appear on stdout: libraries (NCCL prints its version banner there when NCCL_DEBUG=VERSION)
# are redirected to stderr for the lifetime of the process; emit() writes to the saved descriptor.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 480, 854
METRIC = "frames/sec at 480x854 fwd-only, batch 1 per GPU (OSVOS.forward -> 5 logit maps)"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"],
                "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([v.strip() for v in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active")
                                                         for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def cpu_reference_fps(steps, warmup, workload):
    """The reference's own CPU path (PyTorch fp32 / MKLDNN on all host cores), restated by the oracle port."""
    import torch
    from oracle import osvos_oracle as oc
    cores = os.cpu_count() or 1
    params = oc.he_params(seed=0)
    x, gt = oc.synthetic_frame(1, H, W, 1234)

    def probe(threads):
        torch.set_num_threads(threads)
        with torch.no_grad():
            oc.osvos_forward(params, x)          # warm-up (thread pool, MKLDNN primitives)
            t0 = time.perf_counter()
            oc.osvos_forward(params, x)
        return time.perf_counter() - t0
    # "all the host threads it can use": torch's CPU conv slows down when oversubscribed on many-core hosts,
    # so the thread count is the fastest of {all cores, 64, 32, 16} on a one-frame probe.
    cands = sorted({c for c in (cores, 64, 32, 16) if c <= cores}, reverse=True)
    best = min(cands, key=probe)
    torch.set_num_threads(best)

    def one():
        if workload == "train480":
            oc.forward_backward(params, x, gt, objective="online")
        else:
            with torch.no_grad():
                oc.osvos_forward(params, x)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    return 1.0 / dt, dt * 1e3, cores, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 2))
    fps, ms, cores, threads = cpu_reference_fps(steps, warmup, args.workload)
    line = {"impl": "reference", "metric": METRIC if args.workload == "infer480" else METRIC.replace("fwd-only", "fwd+bwd"),
            "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: 1x3x{H}x{W} synthetic BGR frame, OSVOS VGG-16 trunk, He-init weights"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": f"{steps} steps of the full 480x854 frame after {warmup} warm-up, torch CPU fp32 "
                                       f"(MKLDNN) on {threads} threads of {cores} host cores"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_parent(args, rank, world, local, dev):
    """BASELINE.json configs[3]: parent training, per-GPU batch 12 at 480x854, 5-loss objective, one
    gradient allreduce(mean) + SGD step per step; weak scaling (per-GPU work fixed)."""
    import torch
    import torch.distributed as dist
    from osvos_pytorch_b200 import ops, parallel, training
    from osvos_pytorch_b200.networks.vgg_osvos import OSVOS, he_init_
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    net = he_init_(OSVOS(pretrained=0, verbose=False, precision=args.precision), seed=0).to(dev)
    opt = training.make_optimizer(net, "parent", fused=True)   # one-launch SGD + grad zeroing + weight repack
    bucket = parallel.GradientBucket(parallel.trainable_parameters(net), dev)
    batches = [training.synthetic_batch(args.batch, H, W, 1000 * rank + i, dev) for i in range(2)]

    def one(i):
        training.parent_epoch(net, opt, bucket, [batches[i % 2]], 0, 240, 1)
    for i in range(warmup):
        one(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ops.KERNEL_LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        one(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t) / steps
    launches = (ops.KERNEL_LAUNCHES[0] - l0) // steps
    clocks = sampler.stop() if rank == 0 else None
    # allreduce alone (device time), for the share of the step
    ar_ms = 0.0
    if world > 1:
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            bucket.allreduce_mean()
        a1.record()
        torch.cuda.synchronize()
        ar_ms = a0.elapsed_time(a1) / 10
    if rank == 0:
        line = {"metric": "frames/sec at 480x854 fwd+bwd, parent training (5-loss objective, SGD step, DP allreduce)",
                "value": world * args.batch * 1000.0 / ms, "unit": "frames/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16x3 (split-bf16 operands, fp32 accumulate)" if args.precision == "exact" else "bf16",
                "data": "synthetic",
                "config": {"workload": f"parent480: per-GPU batch {args.batch} x 3x{H}x{W} synthetic frames, global batch "
                                       f"{world * args.batch}, parent objective, SGD(lr 1e-8, mom .9, wd 2e-4)",
                           "parallelism": f"dp{world}: allreduce(mean) of a 59.7 MB flat fp32 gradient bucket per step (NCCL)",
                           "l2": "per-step working set (>10 GB) exceeds L2", "timing": "CUDA events, max over ranks"},
                "allreduce_ms": ar_ms, "allreduce_share": ar_ms / ms if ms else None,
                "gpu_launches": int(launches), "clocks": clocks}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


NCU_CONV_DRAM_BYTES_PER_STEP = 465.34e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="infer480", choices=["infer480", "train480", "parent480"])
    ap.add_argument("--batch", type=int, default=12, help="parent480: frames per GPU per optimizer step")
    ap.add_argument("--precision", default="exact", choices=["exact", "fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager-train", action="store_true", help="train480: eager launches instead of the step graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from oracle import osvos_oracle as oc           # cpu_baseline leg + synthetic input generator only
    from osvos_pytorch_b200 import ops
    from osvos_pytorch_b200.layers.osvos_layers import class_balanced_cross_entropy_loss
    from osvos_pytorch_b200.networks.vgg_osvos import OSVOS, he_init_

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    if args.workload == "parent480":
        return run_parent(args, rank, world, local, dev)
    train = args.workload == "train480"

    net = he_init_(OSVOS(pretrained=0, verbose=False, precision=args.precision), seed=0).to(dev)
    net.train(train)
    n_in = 4                                          # rotate input frames (distinct seeds)
    frames = [oc.synthetic_frame(1, H, W, 1234 + i + 100 * rank) for i in range(n_in)]
    xs = [f[0].to(dev) for f in frames]
    gts = [f[1].to(dev) for f in frames]
    xs_host = [f[0].pin_memory() for f in frames]
    out_host = torch.empty((1, 1, H, W), dtype=torch.float32).pin_memory()
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    graphed = {"step": None}

    def step(i, x=None):
        x = xs[i % n_in] if x is None else x
        if train:
            if args.eager_train:
                net.zero_grad(set_to_none=False)
                outs = net(x)
                loss = class_balanced_cross_entropy_loss(outs[-1], gts[i % n_in], size_average=False)
                loss.backward()
                return loss
            # fwd + online loss + bwd of the micro-batch as one replayed CUDA graph (osvos_pytorch_b200.training)
            sample = {"image": x, "gt": gts[i % n_in]}
            if graphed["step"] is None:
                from osvos_pytorch_b200.training import GraphedTrainStep
                graphed["step"] = GraphedTrainStep(
                    net, lambda outs, gt: class_balanced_cross_entropy_loss(outs[-1], gt, size_average=False), sample)
            return graphed["step"](sample)          # gradients accumulate, as between the reference's optimizer steps
        with torch.no_grad():
            return net(x)[-1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / k

    # kernels per step, counted on an eager pass (the timed inference steps replay a captured CUDA graph of
    # exactly these launches; the training step is always eager)
    graphs_on = net._engine.use_cuda_graph
    net._engine.use_cuda_graph = False
    eager_flag = args.eager_train
    args.eager_train = True                          # count launches on an eager pass
    step(0)
    l0 = ops.KERNEL_LAUNCHES[0]
    step(1)
    launches = ops.KERNEL_LAUNCHES[0] - l0
    args.eager_train = eager_flag
    net._engine.use_cuda_graph = graphs_on
    for i in range(warmup):
        step(i)
    # W steps are ~16 ms of work: not enough for the clocks / power state of a box that was idle to settle (the e2e
    # figure, measured seconds later, used to come out FASTER than the device-resident one).  Keep stepping, untimed,
    # for half a second before the timed region.
    t_settle, i = time.perf_counter(), warmup
    while time.perf_counter() - t_settle < 0.5:
        for _ in range(20):
            step(i)
            i += 1
        torch.cuda.synchronize()
    # ---- device-resident throughput -------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step, steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: pinned host frame -> H2D -> OSVOS.forward -> D2H of the result ---
    def e2e_step(i):
        x = xs_host[i % n_in].to(dev, non_blocking=True)
        r = step(i, x)
        if train:
            loss_host.copy_(r.detach(), non_blocking=True)
        else:
            out_host.copy_(r, non_blocking=True)
    e2e_extra = {}
    if train:
        for i in range(3):
            e2e_step(i)
        ms_e2e = timed(e2e_step, steps)
    else:
        # the test-time loop of the reference (train_online.py:172-187) through the package's sequence pipeline:
        # every frame is copied H2D from pinned memory, run through OSVOS.forward, and its result copied D2H -
        # the three legs of consecutive frames overlap on separate streams (osvos_pytorch_b200/inference.py)
        from osvos_pytorch_b200.inference import SequenceSegmenter

        def timed_sequence(seg, k):
            for _ in seg(xs_host[i % n_in] for i in range(6)):      # warm-up, allocates the ring
                pass
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in seg(xs_host[i % n_in] for i in range(k)):
                pass
            seg.join_current_stream()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t) / k
        ms_e2e = timed_sequence(SequenceSegmenter(net, output="logits"), steps)
        for i in range(3):
            e2e_step(i)
        ms_serial = timed(e2e_step, steps)
        ms_png = timed_sequence(SequenceSegmenter(net, output="bytescale"), steps)
        e2e_extra = {"serial_single_stream": {"value": world * 1000.0 / ms_serial, "ms_per_step": ms_serial},
                     "u8_png_payload": {"value": world * 1000.0 / ms_png, "ms_per_step": ms_png,
                                        "d2h_bytes_per_step": H * W,
                                        "note": "sigmoid + imsave bytescale on the device (ops.logits_to_u8)"}}

    # ---- roofline of the dominant kernel (tcgen05 conv): CUDA events around every launch ---
    conv_ms, conv_flops, conv_calls, eager_ms = 0.0, 0.0, 0, 0.0
    if rank == 0 and not train:
        rec = []
        orig = ops.conv3x3

        def wrapped(x, w_packed, bias, cout, *a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig(x, w_packed, bias, cout, *a, **k)
            e.record()
            n_, hh, ww, ci = x.shape
            rec.append((s, e, 2.0 * n_ * hh * ww * cout * 9 * ci))
            return r
        ops.conv3x3 = wrapped
        import osvos_pytorch_b200.engine as eng
        eng.ops.conv3x3 = wrapped
        net._engine.use_cuda_graph = False          # per-launch events need the eager path
        reps = min(steps, 10)
        for i in range(3):                          # eager warm-up passes, not counted
            step(i)

        def instrumented(park_gpu):
            """`reps` eager passes with an event pair around every conv launch.  An eager launch costs the host ~40 us
            (ctypes + six tensor-map encodes), more than the short kernels take, so with the GPU idle the event pairs
            would time the HOST.  park_gpu: a ~40 ms spin kernel is enqueued first and every launch of the passes
            queues up behind it; the GPU then runs them back to back and the events see kernel time only."""
            torch.cuda.synchronize()
            rec.clear()
            if park_gpu:
                torch.cuda._sleep(int(0.04 * 1.9e9))
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for i in range(reps):
                step(i)
            p1.record()
            torch.cuda.synchronize()
            return p0.elapsed_time(p1) / reps       # the instrumented (eager, per-launch events) step
        parked = True
        try:
            eager_ms = instrumented(True)
        except Exception:                            # torch.cuda._sleep is a private helper: fall back to plain eager
            parked = False
            eager_ms = instrumented(False)
        ops.conv3x3 = orig
        eng.ops.conv3x3 = orig
        net._engine.use_cuda_graph = graphs_on
        conv_ms = sum(s.elapsed_time(e) for s, e, _ in rec) / reps
        conv_flops = sum(f for _, _, f in rec) / reps
        conv_calls = len(rec) // reps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    fps = world * 1000.0 / ms
    line = {
        "metric": METRIC if not train else METRIC.replace("fwd-only", "fwd+bwd (online objective)"),
        "value": fps, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (split-bf16 operands, fp32 accumulate)" if args.precision == "exact" else "bf16",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: 1x3x{H}x{W} synthetic BGR frame per step, OSVOS VGG-16 trunk + 4 side "
                               f"branches, He-init weights, precision={args.precision}",
                   "parallelism": f"replicas x{world} (no collective on this path)",
                   "l2": "per-step activation traffic (~0.9 GB exact) exceeds the 126 MB L2; inputs rotate over 4 frames; no explicit flush",
                   "timing": "CUDA events on the launching stream, max over ranks; W warm-up steps + 0.5 s of untimed steps first",
                   "launch": ("captured CUDA graph of the step's kernels, replayed per step"
                              if ((graphs_on and not train) or (train and not args.eager_train)) else "eager launches")},
        "e2e": {"value": world * 1000.0 / ms_e2e, "unit": "frames/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": 3 * H * W * 4, "d2h_bytes_per_step": (4 if train else H * W * 4),
                "path": ("pinned host frame -> .to(cuda) -> fwd+loss+bwd -> D2H of the loss" if train else
                         "SequenceSegmenter: pinned host frame -> H2D -> OSVOS.forward (nn.Module API) -> D2H of the "
                         "fused logit map, legs of consecutive frames overlapped on 3 streams"), **e2e_extra},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if conv_ms > 0:
        ach = conv_flops / (conv_ms * 1e-3) / 1e12
        line["roofline"] = {"bound": "tensor", "kernel": "conv3x3_halo_kernel (tcgen05 implicit GEMM, halo reuse; 16 launches per step)",
                            "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                            "frac": ach / peaks["tflops_sustained"],
                            "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside the step)",
                            "algorithmic_flops_per_step": conv_flops, "launches_per_step": conv_calls,
                            "kernel_ms_per_step": conv_ms,
                            # share measured inside ONE pass: per-launch events and the pass's own total (eager
                            # launches; the headline `ms_per_step` replays the same kernels from a CUDA graph)
                            "share_of_step": conv_ms / eager_ms, "instrumented_step_ms": eager_ms,
                            "instrumented_how": ("eager launches queued behind a parked GPU (back-to-back kernel time)"
                                                 if parked else "plain eager launches (host launch gaps included)"),
                            "tensor_pipe_passes": 3 if args.precision == "exact" else 1,
                            # exact mode emulates fp32 operands with three bf16 passes (hi*hi + hi*lo + lo*hi): the
                            # tensor pipe EXECUTES passes x the algorithmic flops; this is that figure over the peak
                            "issued_mma_frac": ach * (3 if args.precision == "exact" else 1) / peaks["tflops_sustained"],
                            # dram__bytes_read.sum + dram__bytes_write.sum of the 16 conv launches of one 480x854 exact
                            # frame, from the committed ncu launch list (profiles/r01f_launches_infer480.csv; the
                            # `--set full` capture r01d_ncu_full_forward_kernels.csv had 476.2 MB): 465.3 MB per step =
                            # 29.1 MB per launch (activations in + out; weights stay in L2)
                            "traffic": (NCU_CONV_DRAM_BYTES_PER_STEP / conv_calls
                                        if args.precision == "exact" and conv_calls == 16 else None),
                            "traffic_unit": "bytes per launch (average over the step's conv launches)",
                            "traffic_source": "profiles/r01f_launches_infer480.csv"}
    if not args.no_cpu_baseline:
        cfps, cms, cores, threads = cpu_reference_fps(3, 1, args.workload)
        line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": threads, "kind": "port",
                                "sample": f"3 steps of the same 480x854 frame after 1 warm-up; oracle port = the "
                                          f"reference's torch CPU fp32 path on {threads} threads ({cores} host cores)"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
