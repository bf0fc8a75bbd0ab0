#!/usr/bin/env python
"""bench.py -- scene graphs/sec of the SG -> layout training step (GCN + layout, fwd + bwd).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...     # CPU restatement of the reference (oracle port)

Workload (BASELINE.json configs[1], "cfg2"): 128 scene graphs per GPU, 3-30 objects each (+ the
__image__ dummy), VG-like vocabulary (50 predicates), WSGC canonicalization with learned converse +
transitive edges, 5-layer GraphTripleConv stack (embed 128 / hidden 512) + box_net, boxes_to_layout
64x64x128 canvas on the GT boxes, backward through both, Adam.  Synthetic graphs, random-init weights.
One JSON line on rank 0; see the prompt's bench contract for the keys.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

from canonicalsg2im_b200 import synth   # noqa: E402

METRIC = "scene_graphs_per_sec_gcn_layout_fwd_bwd"
UNIT = "graphs/s"
WORKLOAD = ("cfg2: packed_vg-like SG->layout training step with WSGC canonicalization, batch 128/GPU, 3-30 objects, "
            "P=50, 5x GraphTripleConv(128/512) + box_net + boxes_to_layout 64x64x128, fwd+bwd+Adam")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CSG_PRECISION", "bf16"), choices=["fp32", "bf16"],
                    help="bf16 = tcgen05 tensor-core engine (north_star: bf16 MLP, fp32 accumulate, 1e-2 rel); "
                         "fp32 = the 1e-5 parity engine")
    ap.add_argument("--batch", type=int, default=128, help="graphs per GPU")
    ap.add_argument("--cpu-sample", type=int, default=8, help="graphs in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="profiling run: cudaProfilerStart/Stop around the timed region (ncu --profile-from-start off), "
                         "no e2e / CPU legs; the printed numbers are not bench values")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_graphs(batch, seed):
    vocab = synth.Vocab(42)
    return vocab, synth.make_graphs(batch, 1000 + seed, 3, 30, vocab, include_dummies=True)


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_step_runner(vocab, sample, threads):
    from oracle.step import CpuStep
    torch.set_num_threads(threads)
    _, graphs = workload_graphs(sample, 0)
    W = synth.make_conv_weights(vocab, 0)
    step = CpuStep(vocab, synth.make_state(vocab, seed=0), W)
    uni = synth.det_uniform(sum(len(g.triplets) for g in graphs), 13)
    return step, graphs, uni


def run_cpu_baseline(sample, threads, budget_s=25.0):
    vocab = synth.Vocab(42)
    step, graphs, uni = cpu_step_runner(vocab, sample, threads)
    step.step(graphs, uni)                       # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        step.step(graphs, uni)
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 5:
            break
    dt = (time.perf_counter() - t0) / n
    return {"value": sample / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d graphs of the cfg2 workload per step, %d timed steps, oracle/step.py (torch-CPU + numpy)"
                      % (sample, n)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vocab = synth.Vocab(42)
    sample = max(1, args.cpu_sample // 2)         # bounded so that K steps end within a few minutes
    step, graphs, uni = cpu_step_runner(vocab, sample, threads)
    step.step(graphs, uni)                       # one bounded warm-up step (each CPU step costs seconds)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step.step(graphs, uni)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_graphs_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d graphs of the cfg2 workload per step (oracle/step.py, torch-CPU + numpy)" % sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def layout_roofline(d, G, pk, iters=20):
    """The layout compositor pair of this workload timed alone (CUDA events on the launching stream; the 268 MB
    canvas and its gradient exceed the 126 MB L2, so every launch streams from / to HBM).  Algorithmic bytes:
    forward = one write of N*D*H*W*4, backward = one read of the same."""
    from canonicalsg2im_b200 import _lib
    from canonicalsg2im_b200.layout import _linspace
    from canonicalsg2im_b200.ops import lib, ptr, workspace, _stream
    L = lib()
    dev = G.device
    N, D, H, W = G.shape
    boxes, off = d["boxes"].float().contiguous(), d["obj_off"]
    NO = boxes.shape[0]
    vecs = torch.randn((NO, D), device=dev)
    out = torch.empty_like(G)
    dv = torch.empty((NO, D), device=dev)
    lx, ly = _linspace(W, dev), _linspace(H, dev)
    ws = workspace(L.csg_layout_bwd_vecs_workspace(N, NO, D, H, W), dev)
    mo = int(d["max_objs"])

    def fwd():
        _lib.check(L.csg_layout_fwd(ptr(vecs), ptr(boxes), 0, ptr(off), ptr(lx), ptr(ly), ptr(out), N, D, H, W, 0, 0, mo,
                                    _stream()), "csg_layout_fwd")

    def bwd():
        _lib.check(L.csg_layout_bwd_vecs(ptr(G), ptr(boxes), 0, ptr(off), ptr(lx), ptr(ly), ptr(dv), N, NO, D, H, W, 0, 0,
                                         mo, ptr(ws), ws.numel(), _stream()), "csg_layout_bwd_vecs")
    res = {}
    nbytes = N * D * H * W * 4
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3 / iters
        res[name] = {"us": sec * 1e6, "achieved": nbytes / sec / 1e9, "frac": nbytes / sec / 1e9 / pk["hbm"]}
    return {"kernel": "layout_fwd_kernel / layout_bwd_ring_kernel (boxes_to_layout %dx%dx%dx%d)" % (N, D, H, W),
            "bound": "hbm", "unit": "GB/s", "peak": pk["hbm"], "peak_source": pk["src"] + " copy bandwidth",
            "bytes_per_launch": nbytes, "fwd": res["fwd"], "bwd": res["bwd"]}


def run_ours(args):
    import torch.distributed as dist
    from canonicalsg2im_b200 import _lib, ops
    from canonicalsg2im_b200.pipeline import SgToLayoutStep, HostBatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device for --impl ours (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.load()
    pk = peaks()

    vocab, graphs = workload_graphs(args.batch, rank)
    hb = HostBatch(graphs, seed=rank)
    step = SgToLayoutStep(vocab, dev, precision=args.precision, distributed=world > 1, seed=0)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    G = torch.randn((args.batch, 128, 64, 64), device=dev, generator=gen) * 1e-3
    d = hb.to_device(dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks / throttle reasons are sampled every 20 ms from the warm-up to the end of the end-to-end region (all of it
    # is the same load; the device-resident timed region alone lasts ~0.1 s)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n_tri = 0
    for _ in range(max(args.warmup, 3)):
        _, n_tri = step.step(d, G, prefetch=d)
    barrier()

    # ---- timed region 1: inputs resident in HBM
    launches0 = L.csg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.profile:
        torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    for _ in range(args.steps):
        # every step canonicalizes its own batch; the COUNTING pass of the next step's batch is enqueued one step
        # ahead (as the reference's DataLoader workers do), so K steps contain exactly K count + K emit passes
        loss, n_tri = step.step(d, G, prefetch=d)
    e1.record()
    barrier()
    if args.profile:
        torch.cuda.cudart().cudaProfilerStop()
    launches = L.csg_launch_count() - launches0
    sec = e0.elapsed_time(e1) * 1e-3
    # ---- instrumented pass (not the headline): the same K steps with CUDA events around every GEMM / layout /
    # pooling entry point on the launching stream (csg_prof_*), for the roofline objects
    import ctypes
    prof = (ctypes.c_double * 24)()
    L.csg_prof_enable(1)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(0 if args.profile else args.steps):
        step.step(d, G, prefetch=d)
    p1.record()
    torch.cuda.synchronize()
    L.csg_prof_collect(prof)
    L.csg_prof_enable(0)
    prof_sec = p0.elapsed_time(p1) * 1e-3
    cls = 1 if args.precision == "fp32" else 0
    gemm_flops, gemm_sec, gemm_n = prof[cls * 3], prof[cls * 3 + 1], int(prof[cls * 3 + 2])
    tsec = torch.tensor([sec], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tsec, op=dist.ReduceOp.MAX)
    sec = float(tsec.item())
    value = world * args.batch * args.steps / sec

    # ---- timed region 2: end to end through the public API with HOST buffers (H2D + D2H every step)
    d.pop("_canon_plan", None)
    dd = hb.to_device(dev)
    for _ in range(0 if args.profile else 2):
        nxt = hb.to_device(dev)
        l, _ = step.step(dd, G, prefetch=nxt)
        dd = nxt
        float(l.item())
    barrier()
    t0 = time.perf_counter()
    lv = float("nan")
    loss_host = torch.empty(2, dtype=torch.float32, pin_memory=True)
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    for i in range(0 if args.profile else args.steps):
        # host buffers -> device for the NEXT step and its canonicalization counting pass are enqueued while this
        # step runs (one upload + one count + one emit per step inside the timed region).  Every step's loss is copied
        # to pinned host memory; the host reads it one step later (as a logging loop would), so that the read never
        # drains the launch queue.  The last loss is read before the clock stops.
        nxt = hb.to_device(dev)
        l, _ = step.step(dd, G, prefetch=nxt)
        dd = nxt
        loss_host[i & 1].copy_(l, non_blocking=True)
        loss_ev[i & 1].record()
        if i > 0:
            loss_ev[(i - 1) & 1].synchronize()
            lv = float(loss_host[(i - 1) & 1])
    if not args.profile and args.steps > 0:
        loss_ev[(args.steps - 1) & 1].synchronize()
        lv = float(loss_host[(args.steps - 1) & 1])
    torch.cuda.synchronize()
    e2e_sec = time.perf_counter() - t0
    tsec = torch.tensor([e2e_sec], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tsec, op=dist.ReduceOp.MAX)
    e2e_sec = float(tsec.item())
    clocks = sampler.stop() if rank == 0 else None
    e2e = {"value": world * args.batch * args.steps / max(e2e_sec, 1e-9), "unit": UNIT,
           "h2d_bytes_per_step": int(hb.nbytes), "d2h_bytes_per_step": 4 + 8,   # loss scalar + the two canonicalization size words
           "ms_per_step": 1e3 * e2e_sec / args.steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm = None if args.profile else layout_roofline(d, G, pk)
    n_obj = int(hb.obj_off[-1])
    peak_tf = pk["tf_sustained"]
    ach_tf = gemm_flops / gemm_sec / 1e12 if gemm_sec > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.precision)
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "graphs_per_gpu": args.batch, "objects": n_obj, "triples_after_canon": n_tri,
                   "precision": args.precision, "parallelism": "graph-sharded dp%d" % world,
                   "l2": "per-step working set (net1 activations %.0f MB + 268 MB canvas + 268 MB canvas grad) exceeds the 126 MB L2"
                         % (n_tri * (1152 + 512) * (4 if args.precision == "fp32" else 2) / 1e6)},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {"kernel": "net1/net2 GEMMs (%s)" % ("gemm_f32_kernel" if args.precision == "fp32" else "gemm_tc_kernel"),
                     "bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": ach_tf / peak_tf, "traffic": traffic, "peak_source": pk["src"] + " bf16 sustained",
                     "launches_timed": gemm_n, "share_of_step": gemm_sec / prof_sec if prof_sec > 0 else None,
                     "note": "measured in a separate instrumented pass of the same %d steps (%.3f ms/step)"
                             % (args.steps, 1e3 * prof_sec / args.steps)},
        "roofline_hbm": hbm,
        "loss": lv,
    }
    if not args.no_cpu_baseline and not args.profile and world == 1:
        try:
            line["cpu_baseline"] = run_cpu_baseline(args.cpu_sample, os.cpu_count() or 1)
        except Exception as ex:   # the baseline is a reported number, never a reason to lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": "failed: %r" % (ex,)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
