"""Device timeline of one training step (torch.profiler / CUPTI): kernel busy time, idle gaps and where they are.
    python scratch/timeline.py [--precision bf16] > gpurun_out/timeline.txt"""
import os
import sys
import json
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
from canonicalsg2im_b200 import synth  # noqa: E402
from canonicalsg2im_b200.pipeline import SgToLayoutStep, HostBatch  # noqa: E402

prec = "bf16" if "--precision" not in sys.argv else sys.argv[sys.argv.index("--precision") + 1]
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
vocab = synth.Vocab(42)
graphs = synth.make_graphs(128, 1000, 3, 30, vocab, include_dummies=True)
hb = HostBatch(graphs, seed=0)
step = SgToLayoutStep(vocab, dev, precision=prec, seed=0, use_graph="--graph" in sys.argv)     # --graph: CUDA-graph replay
G = torch.randn((128, 128, 64, 64), device=dev) * 1e-3
d = hb.to_device(dev)
for _ in range(5):
    step.step(d, G, prefetch=d)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step.step(d, G, prefetch=d)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace.json")
prof.export_chrome_trace(path)
tr = json.load(open(path))
ev = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
busy = sum(e["dur"] for e in ev)
print("3 steps: span %.1f us, kernel busy %.1f us (%.1f%%), %d device ops" % (t1 - t0, busy, 100 * busy / (t1 - t0), len(ev)))
gaps = []
end = ev[0]["ts"] + ev[0]["dur"]
for a, b in zip(ev[:-1], ev[1:]):
    end = max(end, a["ts"] + a["dur"])
    g = b["ts"] - end
    if g > 3:
        gaps.append((g, a["name"][:60], b["name"][:60], b["ts"] - t0))
print("gaps > 3us: %d, total %.1f us" % (len(gaps), sum(g[0] for g in gaps)))
for g in sorted(gaps, key=lambda g: -g[0])[:40]:
    print("  %8.1f us at %9.1f  after %-60s before %s" % (g[0], g[3], g[1], g[2]))
# sequential listing of the middle step
n = len(ev) // 3
print("---- ops of the middle step (ts rel, dur, gap-before, name)")
prev_end = None
for e in ev[n:2 * n]:
    gap = 0 if prev_end is None else e["ts"] - prev_end
    print("%9.1f %8.1f %7.1f  %s" % (e["ts"] - t0, e["dur"], gap, e["name"][:90]))
    prev_end = max(prev_end or 0, e["ts"] + e["dur"])
# CPU-side: top host ops by self time
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25))
