"""Launches the three STDiT3-XL/2 attention shapes once each (for ncu) and times them with CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-sora_b200"))
import torch

import osb200 as osb

osb.init(0)
B, T, S, H, D, Ly = 1, 64, 256, 16, 72, 300
C, N = H * D, T * S
qkv = torch.randn(B * N, 3 * C, device="cuda").bfloat16()
qc = torch.randn(B * N, C, device="cuda").bfloat16()
kv = torch.randn(B * Ly, 2 * C, device="cuda").bfloat16()
out = torch.empty(B * N, C, device="cuda", dtype=torch.bfloat16)
w = torch.ones(D, device="cuda").bfloat16()
inv = 1.0 / (10000 ** (torch.arange(0, D, 2, device="cuda").float() / D))
ang = torch.arange(T, device="cuda").float()[:, None] * inv[None]
cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
lens = torch.tensor([260], device="cuda", dtype=torch.int32)


def spatial():
    osb.attn_short(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, num_seqs=B * T, seqs_per_batch=T, q_strides=(N, S, 1),
                   k_strides=(N, S, 1), Lq=S, Lk=S, num_heads=H, head_dim=D, q_norm_w=w, k_norm_w=w)


def temporal():
    osb.attn_short(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, num_seqs=B * S, seqs_per_batch=S, q_strides=(N, 1, S),
                   k_strides=(N, 1, S), Lq=T, Lk=T, num_heads=H, head_dim=D, q_norm_w=w, k_norm_w=w, rope_cos=cos, rope_sin=sin)


def cross():
    osb.attn_short(qc, kv[:, :C], kv[:, C:], out, num_seqs=B, seqs_per_batch=1, q_strides=(N, 0, 1), k_strides=(Ly, 0, 1),
                   Lq=N, Lk=Ly, num_heads=H, head_dim=D, kv_lens=lens)


for impl in ([int(v) for v in sys.argv[1:]] or [0]):
  osb.ATTN_IMPL = impl
  for name, fn, flops in (("spatial", spatial, 4 * N * S * C), ("temporal", temporal, 4 * N * T * C), ("cross", cross, 4 * N * Ly * C)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"attn impl{impl} {name:9s}: {ms*1e3:8.1f} us  {flops/(ms*1e-3)/1e12:6.1f} TF/s")
