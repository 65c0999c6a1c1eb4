"""Host-side profile of one STDiT3-XL/2 forward (debug aid): where does the Python side spend its time?"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "open-sora_b200"))
import torch

import bench

dev = torch.device("cuda", 0)
import osb200

osb200.init(0)
model = bench.build_model(dev)
hin = bench.host_inputs(1)
din = {k: v.to(dev) for k, v in hin.items()}
with torch.no_grad():
    for _ in range(3):
        model(**din)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model(**din)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host issue time {1e3 * (t1 - t0):.1f} ms, until GPU done {1e3 * (t2 - t0):.1f} ms")
    pr = cProfile.Profile()
    pr.enable()
    model(**din)
    pr.disable()
    torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
print("tmap cache hits/misses", osb200.tmap_cache_stats())


def loop(tag, inp, n=10):
    with torch.no_grad():
        for _ in range(2):
            model(**inp)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        for _ in range(n):
            model(**inp)
        e.record()
        torch.cuda.synchronize()
        print(f"{tag}: {s.elapsed_time(e) / n:.2f} ms/step (events), {1e3 * (time.perf_counter() - t0) / n:.2f} ms wall", flush=True)


loop("device height/width (sync per forward)", din)
din2 = dict(din, height=hin["height"], width=hin["width"])
loop("host height/width", din2)
cs = bench.ClockSampler(0)
cs.start()
loop("host height/width + nvidia-smi sampler", din2)
print(cs.stop())
os.environ["OSB_ATTN_TILES"] = "0"
loop("register-path attention", din2)

# ---- where does a synchronised forward spend its time?  per-family device time under each condition
os.environ["OSB_ATTN_TILES"] = "1"
for tag, inp in (("sync", din), ("nosync", din2)):
    with torch.no_grad():
        for _ in range(2):
            model(**inp)
        osb200.start_profile()
        model(**inp)
        rec = osb200.stop_profile()
    fam = {}
    for name, work, t in rec:
        f = fam.setdefault(name, [0, 0.0])
        f[0] += 1
        f[1] += t
    print(tag, {k: (v[0], round(v[1], 2)) for k, v in fam.items()}, flush=True)
