"""GEMM tile-shape sweep on the STDiT3-XL/2 shapes (device timing with CUDA events, L2 flushed by
rotating over operand sets larger than L2).  Prints TFLOP/s per (shape, block_n, cta_group)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-sora_b200"))
import torch

import osb200 as osb

osb.init(0)
M = 16384
SHAPES = [("qkv", 3456, 1152, osb.EPI_BIAS), ("proj+gate+res", 1152, 1152, osb.EPI_BIAS_GATE_RES),
          ("fc1+gelu", 4608, 1152, osb.EPI_BIAS_GELU_TANH), ("fc2+gate+res", 1152, 4608, osb.EPI_BIAS_GATE_RES)]
NSETS = 4
for name, N, K, epi in SHAPES:
    sets = []
    for i in range(NSETS):
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        r = torch.randn(M, N, device="cuda").bfloat16()
        o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        sets.append((a, w, b, r, o))
    gate = torch.randn(1, N, device="cuda")
    ref = torch.matmul(sets[0][0], sets[0][1].t())
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(20):
        torch.matmul(sets[i % NSETS][0], sets[i % NSETS][1].t())
    e.record()
    torch.cuda.synchronize()
    print(f"{name:16s} N={N} K={K}  cuBLAS(no epilogue) {2*M*N*K/ (s.elapsed_time(e)/20*1e-3)/1e12:7.1f} TF/s")
    for cta in (1, 2):
        for bn in (128, 192, 256):
            kw = dict(epilogue=epi, cta_group=cta, block_n=bn)
            if epi == osb.EPI_BIAS_GATE_RES:
                kw.update(gate=gate, group_rows=M)
            def run(i):
                a, w, b, r, o = sets[i % NSETS]
                if epi == osb.EPI_BIAS_GATE_RES:
                    osb.gemm(a, w, b, residual=r, out=o, **kw)
                else:
                    osb.gemm(a, w, b, out=o, **kw)
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            s.record()
            for i in range(20):
                run(i)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 20
            print(f"   cta{cta} bn{bn}: {ms*1e3:8.1f} us  {2*M*N*K/(ms*1e-3)/1e12:7.1f} TF/s")
