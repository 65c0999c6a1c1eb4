"""Back-to-back launches of one kernel at a time (debug aid): which kernel shows rare long stalls?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-sora_b200"))
import torch

import osb200 as osb

osb.init(0)
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
out2 = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
nw = torch.ones(D, device=dev).bfloat16()
inv = 1.0 / (10000 ** (torch.arange(0, D, 2, device=dev).float() / D))
ang = torch.arange(T, device=dev).float()[:, None] * inv[None]
cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
lens = torch.tensor([260], device=dev, dtype=torch.int32)
qkv = torch.empty(B * N, 3 * C, device=dev, dtype=torch.bfloat16)
zsh = torch.zeros(1, C, device=dev)
sp_t = osb.HeadTiles(B * N, osb.tile_map(0, S), 3, H, D, dev)
tm_t = osb.HeadTiles(B * N, osb.tile_map(0, T), 3, H, D, dev)
q_t = osb.HeadTiles(B * N, osb.tile_map(0, N, pack=False), 1, H, D, dev)
kv_t = osb.HeadTiles(B * Ly, osb.tile_map(0, Ly, keys_only=True), 2, H, D, dev)
osb.gemm_head_tiles(y, wkv, None, kv_t, nkinds=2)
tm_out = osb.tile_map(1, T, S, T)
cases = {
    "gemm_ht spatial": lambda: osb.gemm_head_tiles(x, wqkv, bqkv, sp_t, nkinds=3, norm_w=(nw, nw, None)),
    "gemm_ht temporal": lambda: osb.gemm_head_tiles(x, wqkv, bqkv, tm_t, nkinds=3, norm_w=(nw, nw, None), rope=(cos, sin), rope_kinds=3),
    "gemm_ht q": lambda: osb.gemm_head_tiles(x, wq, None, q_t, nkinds=1),
    "attn spatial": lambda: osb.attn_tiles(sp_t, sp_t, out, Lk=S, num_seqs=B * T),
    "attn temporal": lambda: osb.attn_tiles(tm_t, tm_t, out, Lk=T, num_seqs=B * S, out_map=tm_out),
    "attn cross": lambda: osb.attn_tiles(q_t, kv_t, out, q_kind=0, k_kind=0, v_kind=1, Lk=Ly, num_seqs=B, kv_lens=lens),
    "ln transposing": lambda: osb.ln_modulate(x, zsh, zsh, group_rows=B * N, scatter=osb.make_scatter(3, 1, 0, T, S, [out2])),
    "gemm plain": lambda: osb.gemm(x, wqkv, bqkv, out=qkv),
}
seq = ["ln transposing", "gemm_ht temporal", "attn temporal", "gemm_ht q", "attn cross", "gemm_ht spatial", "attn spatial"]
cases["block sequence"] = lambda: [cases[k]() for k in seq]
only = os.environ.get("ONLY")
for name, fn in cases.items():
    if only and only not in name:
        continue
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    chunks, per = 60, (10 if name == "block sequence" else 50)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(chunks + 1)]
    ev[0].record()
    for c in range(chunks):
        for _ in range(per):
            fn()
        ev[c + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(chunks))
    print(f"{name:18s}: per-chunk ({per} calls) median {ts[chunks // 2]:.2f} ms, max {ts[-1]:.2f}, 2nd {ts[-2]:.2f}", flush=True)
