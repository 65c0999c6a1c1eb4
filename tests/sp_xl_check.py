"""Sequence-parallel parity at the benchmark latent with the XL/2 block shape (depth 2 to keep it short): both exchanges
against the single-GPU forward of the same model.  Run under torchrun; exits hard to avoid teardown hangs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "open-sora_b200"))
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import bench
from opensora.registry import MODELS, build_module

torch.manual_seed(1234)
with torch.device(dev):
    m = build_module(dict(type="STDiT3-XL/2", depth=int(os.environ.get("DEPTH", "2"))), MODELS).eval()
g = torch.Generator(device=dev).manual_seed(1234)
with torch.no_grad():   # identical parameters AND buffers on every rank (see bench.build_model)
    for n, p in m.named_parameters():
        p.copy_(torch.randn(p.shape, generator=g, device=dev) * ((0.7 / (p[0].numel() ** 0.5)) if p.dim() >= 2 else 0.05) + (1.0 if "norm" in n else 0.0))
    for n, b in m.named_buffers():
        if b.is_floating_point():
            b.copy_(torch.randn(b.shape, generator=g, device=dev) / (b.shape[-1] ** 0.5))
m = m.to(device=dev, dtype=torch.bfloat16)
hin = bench.host_inputs(4321)
din = {k: (v if k in ("height", "width") else v.to(dev)) for k, v in hin.items()}
with torch.no_grad():
    single = m(**din).clone()
    for exchange in os.environ.get("EXCHANGES", "nccl,peer").split(","):
        m.enable_sequence_parallel(dist.group.WORLD, exchange=exchange)
        outs = [m(**din).clone() for _ in range(3)]
        kind = m.sp_exchange_kind
        m.enable_sequence_parallel(None)
        torch.cuda.synchronize()
        errs = [float((o.double() - single.double()).norm() / single.double().norm()) for o in outs]
        print(f"[sp{world} rank {rank}] exchange={exchange} ({kind[:30]}) rel_l2 per step {['%.2e' % e for e in errs]} "
              f"bit_identical={[bool(torch.equal(o, single)) for o in outs]}", flush=True)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
