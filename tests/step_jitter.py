"""Per-step device time of consecutive STDiT3-XL/2 forwards (debug aid): is the step time stable, and when it is not, is
a fixed probe kernel (one plain GEMM) slow at the same time (= clocks / power, not the step's kernels)?"""
import os
import subprocess
import sys
import threading

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
a = torch.randn(16384, 1152, device=dev).bfloat16()
w = torch.randn(4608, 1152, device=dev).bfloat16()
o = torch.empty(16384, 4608, device=dev, dtype=torch.bfloat16)
lines = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap,"
                         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.hw_power_brake_slowdown", "--format=csv,noheader", "-i", "0", "-lms", "20"],
                        stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: lines.extend(proc.stdout), daemon=True).start()
import gc

for tiles in (os.environ.get("JIT_CASES", "1,1,1,1").split(",")):
    os.environ["OSB_ATTN_TILES"] = tiles[0]
    if tiles.endswith("nogc"):
        gc.disable()
    else:
        gc.enable()
    with torch.no_grad():
        for _ in range(3):
            model(**din)
        torch.cuda.synchronize()
        n = 60
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * n + 1)]
        ev[0].record()
        import time as _t
        host = []
        for i in range(n):
            h0 = _t.perf_counter()
            model(**din)
            host.append(1e3 * (_t.perf_counter() - h0))
            ev[2 * i + 1].record()
            osb200.gemm(a, w, out=o)
            ev[2 * i + 2].record()
        torch.cuda.synchronize()
    ts = [ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(n)]
    print("tiles=" + tiles, "median %.1f" % sorted(ts)[n // 2], "host median %.1f" % sorted(host)[n // 2], "outliers(>60ms) (step, gpu ms, host ms of steps k-1..k+1):",
          [(i, round(t), [round(h) for h in host[max(i - 1, 0):i + 2]]) for i, t in enumerate(ts) if t > 60], flush=True)
proc.terminate()
clk = [l.strip() for l in lines]
print("nvidia-smi samples:", len(clk))
print("min clock lines:", sorted(clk, key=lambda l: float(l.split(",")[0].split()[0]))[:6])
print("any slowdown:", [l for l in clk if "Active" in l.split(",", 2)[2].replace("Not Active", "")][:6])
