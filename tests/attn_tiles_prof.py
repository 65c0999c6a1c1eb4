"""Times the head-tile path on the three STDiT3-XL/2 attention shapes (for ncu and CUDA-event timing): the projection
GEMM with the head-tile epilogue and the bulk-copy attention kernel, next to the plain projection GEMM + register-path
kernel they replace.  usage: python tests/attn_tiles_prof.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-sora_b200"))
import torch

import osb200 as osb

osb.init(0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B, T, S, H, D, Ly = 1, 64, 256, 16, 72, 300
C, N = H * D, T * S
dev = "cuda"
x = torch.randn(B * N, C, device=dev).bfloat16()
wqkv = (torch.randn(3 * C, C, device=dev) / C**0.5).bfloat16()
bqkv = torch.randn(3 * C, device=dev).bfloat16()
wq = (torch.randn(C, C, device=dev) / C**0.5).bfloat16()
y = torch.randn(B * Ly, C, device=dev).bfloat16()
wkv = (torch.randn(2 * C, C, device=dev) / C**0.5).bfloat16()
out = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
nw = torch.ones(D, device=dev).bfloat16()
inv = 1.0 / (10000 ** (torch.arange(0, D, 2, device=dev).float() / D))
ang = torch.arange(T, device=dev).float()[:, None] * inv[None]
cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
lens = torch.tensor([260], device=dev, dtype=torch.int32)
qkv = torch.empty(B * N, 3 * C, device=dev, dtype=torch.bfloat16)
qc = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
kv = osb.gemm(y, wkv)

sp_t = osb.HeadTiles(B * N, osb.tile_map(0, S), 3, H, D, dev)
tm_t = osb.HeadTiles(B * N, osb.tile_map(1, T, S, T), 3, H, D, dev)
q_t = osb.HeadTiles(B * N, osb.tile_map(0, N, pack=False), 1, H, D, dev)
kv_t = osb.HeadTiles(B * Ly, osb.tile_map(0, Ly, keys_only=True), 2, H, D, dev)
tmT_t = osb.HeadTiles(B * N, osb.tile_map(0, T), 3, H, D, dev)
zsh = torch.zeros(1, C, device=dev)
osb.gemm_head_tiles(y, wkv, None, kv_t, nkinds=2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2: every timed launch starts cold


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / reps * 1e3


cases = [
    ("qkv gemm -> tiles (spatial)", lambda: osb.gemm_head_tiles(x, wqkv, bqkv, sp_t, nkinds=3, norm_w=(nw, nw, None)), 2.0 * N * 3 * C * C),
    ("qkv gemm -> tiles (temporal)", lambda: osb.gemm_head_tiles(x, wqkv, bqkv, tm_t, nkinds=3, norm_w=(nw, nw, None), rope=(cos, sin), rope_kinds=3), 2.0 * N * 3 * C * C),
    ("qkv gemm -> tiles (temporal, no rope)", lambda: osb.gemm_head_tiles(x, wqkv, bqkv, tm_t, nkinds=3, norm_w=(nw, nw, None)), 2.0 * N * 3 * C * C),
    ("qkv gemm -> tiles (temporal, no rope, no norm)", lambda: osb.gemm_head_tiles(x, wqkv, bqkv, tm_t, nkinds=3), 2.0 * N * 3 * C * C),
    ("qkv gemm -> tiles (temporal, general)", lambda: osb.gemm_head_tiles(x, wqkv, bqkv, tm_t, nkinds=3, norm_w=(nw, nw, None), rope=(cos, sin), rope_kinds=3, general=True), 2.0 * N * 3 * C * C),
    ("qkv gemm -> tiles (spatial, no norm)", lambda: osb.gemm_head_tiles(x, wqkv, bqkv, sp_t, nkinds=3), 2.0 * N * 3 * C * C),
    ("qkv gemm -> tiles (temporal, transposed rows)", lambda: osb.gemm_head_tiles(x, wqkv, bqkv, tmT_t, nkinds=3, norm_w=(nw, nw, None), rope=(cos, sin), rope_kinds=3), 2.0 * N * 3 * C * C),
    ("ln_modulate", lambda: osb.ln_modulate(x, zsh, zsh, group_rows=B * N, out=qc), 4.0 * N * C * 250),
    ("ln_modulate transposing", lambda: osb.ln_modulate(x, zsh, zsh, group_rows=B * N, scatter=osb.make_scatter(3, 1, 0, T, S, [qc])), 4.0 * N * C * 250),
    ("qkv gemm plain", lambda: osb.gemm(x, wqkv, bqkv, out=qkv), 2.0 * N * 3 * C * C),
    ("q gemm -> tiles (cross)", lambda: osb.gemm_head_tiles(x, wq, None, q_t, nkinds=1), 2.0 * N * C * C),
    ("q gemm plain", lambda: osb.gemm(x, wq, out=qc), 2.0 * N * C * C),
    ("attn tiles spatial", lambda: osb.attn_tiles(sp_t, sp_t, out, Lk=S, num_seqs=B * T), 4.0 * N * S * C),
    ("attn tiles temporal", lambda: osb.attn_tiles(tm_t, tm_t, out, Lk=T, num_seqs=B * S), 4.0 * N * T * C),
    ("attn tiles cross", lambda: osb.attn_tiles(q_t, kv_t, out, q_kind=0, k_kind=0, v_kind=1, Lk=Ly, num_seqs=B, kv_lens=lens), 4.0 * N * 260 * C),
    ("attn short spatial", lambda: osb.attn_short(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, num_seqs=B * T, seqs_per_batch=T, q_strides=(N, S, 1), k_strides=(N, S, 1), Lq=S, Lk=S, num_heads=H, head_dim=D, q_norm_w=nw, k_norm_w=nw), 4.0 * N * S * C),
    ("attn short temporal", lambda: osb.attn_short(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, num_seqs=B * S, seqs_per_batch=S, q_strides=(N, 1, S), k_strides=(N, 1, S), Lq=T, Lk=T, num_heads=H, head_dim=D, q_norm_w=nw, k_norm_w=nw, rope_cos=cos, rope_sin=sin), 4.0 * N * T * C),
    ("attn short cross", lambda: osb.attn_short(qc, kv[:, :C], kv[:, C:], out, num_seqs=B, seqs_per_batch=1, q_strides=(N, 0, 1), k_strides=(Ly, 0, 1), Lq=N, Lk=Ly, num_heads=H, head_dim=D, kv_lens=lens), 4.0 * N * 260 * C),
]
only = os.environ.get("ONLY")
for name, fn, flops in cases:
    if only and only not in name:
        continue
    us = timed(fn)
    print(f"{name:32s}: {us:8.1f} us  {flops / us / 1e6:7.1f} TF/s", flush=True)
