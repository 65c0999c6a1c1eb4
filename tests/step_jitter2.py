"""Which launch stalls?  Per-launch device times (CUDA events around every osb200 launch) over many forwards."""
import os
import sys

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
with torch.no_grad():
    for _ in range(3):
        model(**din)
    for rep in range(6):
        osb200.start_profile()
        for _ in range(20):
            model(**din)
        rec = osb200.stop_profile()
        per = len(rec) // 20
        slow = sorted(((t, i, n, w) for i, (n, w, t) in enumerate(rec) if t > 1.0), reverse=True)[:8]
        print(f"rep {rep}: {len(rec)} launches, sum {sum(t for _, _, t in rec) / 20:.1f} ms/step, slow launches:",
              [(round(t, 1), f"step{i // per}", f"launch{i % per}", n, f"{w:.3g}") for t, i, n, w in slow], flush=True)
