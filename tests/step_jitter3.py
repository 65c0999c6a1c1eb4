"""Which HOST call stalls?  Wrap every binding call and torch.empty / torch.zeros with a wall-clock timer during
back-to-back forwards and report calls that took more than 20 ms (debug aid)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "open-sora_b200"))
import torch

import bench
import osb200

dev = torch.device("cuda", 0)
osb200.init(0)
model = bench.build_model(dev)
hin = bench.host_inputs(1)
din = {k: v.to(dev) for k, v in hin.items()}
din.update(height=hin["height"], width=hin["width"])
slow = []
stats = {}


def wrap(mod, name):
    f = getattr(mod, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        dt = 1e3 * (time.perf_counter() - t0)
        s = stats.setdefault(name, [0, 0.0, 0.0])
        s[0] += 1
        s[1] += dt
        s[2] = max(s[2], dt)
        if dt > 20:
            slow.append((name, round(dt, 1), step))
        return r

    setattr(mod, name, g)


for n in ("gemm", "gemm_head_tiles", "attn_tiles", "ln_modulate", "make_scatter", "tile_map"):
    wrap(osb200, n)
for n in ("empty", "zeros", "cat", "empty_like"):
    wrap(torch, n)
step = -1
print({k: v for k, v in os.environ.items() if "PYTORCH" in k or "CUDA" in k})


def mstats(tag):
    st = torch.cuda.memory_stats()
    print(tag, {k: st[k] for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "num_ooms", "reserved_bytes.all.current",
                                   "allocated_bytes.all.current", "segment.all.current", "inactive_split.all.current")}, flush=True)


with torch.no_grad():
    for _ in range(3):
        model(**din)
    torch.cuda.synchronize()
    mstats("before loop")
    for step in range(int(os.environ.get("STEPS", "150"))):
        t0 = time.perf_counter()
        model(**din)
        dt = 1e3 * (time.perf_counter() - t0)
        if dt > 80:
            print(f"step {step}: host forward {dt:.0f} ms; slow calls so far: {slow[-6:]}", flush=True)
            mstats("   ")
    torch.cuda.synchronize()
    mstats("after loop")
print({k: (v[0], round(v[1] / v[0], 3), round(v[2], 1)) for k, v in stats.items()})
