"""Launch each STDiT3-XL/2 GEMM shape (library-chosen tile config) three times over rotating operand sets; run under
`ncu --set full -k regex:gemm_bf16` to capture the current kernel (third launch of each shape = warm)."""
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
for name, N, K, epi in SHAPES:
    sets = []
    for i in range(3):
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        r = torch.randn(M, N, device="cuda").bfloat16()
        o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        sets.append((a, w, b, r, o))
    gate = torch.randn(1, N, device="cuda")
    torch.cuda.synchronize()
    for a, w, b, r, o in sets:
        if epi == osb.EPI_BIAS_GATE_RES:
            osb.gemm(a, w, b, residual=r, out=o, epilogue=epi, gate=gate, group_rows=M)
        else:
            osb.gemm(a, w, b, out=o, epilogue=epi)
    torch.cuda.synchronize()
    print(name, "done")
