"""Timeline of CTA 0 of attn_tiles_kernel (instrumented build, tools/build_trace.sh): per STDiT3-XL/2 attention shape,
the (role, tag, microseconds) events of the first jobs.  Debug aid, not a test."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["OSB200_LIB"] = os.path.join(ROOT, "open-sora_b200", "osb200", "libosb200_trace.so")
sys.path.insert(0, os.path.join(ROOT, "open-sora_b200"))
import torch

import osb200 as osb

osb.init(0)
lib = C.CDLL(os.environ["OSB200_LIB"])
B, T, S, H, D, Ly = 1, 64, 256, 16, 72, 300
Cc, N = H * D, T * S
dev = "cuda"
x = torch.randn(B * N, Cc, device=dev).bfloat16()
wqkv = (torch.randn(3 * Cc, Cc, device=dev) / Cc**0.5).bfloat16()
wq = (torch.randn(Cc, Cc, device=dev) / Cc**0.5).bfloat16()
y = torch.randn(B * Ly, Cc, device=dev).bfloat16()
wkv = (torch.randn(2 * Cc, Cc, device=dev) / Cc**0.5).bfloat16()
out = torch.empty(B * N, Cc, device=dev, dtype=torch.bfloat16)
lens = torch.tensor([260], device=dev, dtype=torch.int32)
sp_t = osb.HeadTiles(B * N, osb.tile_map(0, S), 3, H, D, dev)
tm_t = osb.HeadTiles(B * N, osb.tile_map(1, T, S, T), 3, H, D, dev)
q_t = osb.HeadTiles(B * N, osb.tile_map(0, N, pack=False), 1, H, D, dev)
kv_t = osb.HeadTiles(B * Ly, osb.tile_map(0, Ly, keys_only=True), 2, H, D, dev)
osb.gemm_head_tiles(x, wqkv, None, sp_t, nkinds=3)
osb.gemm_head_tiles(x, wqkv, None, tm_t, nkinds=3)
osb.gemm_head_tiles(x, wq, None, q_t, nkinds=1)
osb.gemm_head_tiles(y, wkv, None, kv_t, nkinds=2)

NAMES = {10: "s_full", 11: "S in regs", 12: "max done", 13: "P stored+arrived", 14: "o_full", 15: "O in regs", 16: "out stored", 17: "stage free", 18: "staged",
         20: "Q0 load", 21: "Q1 load", 22: "KV load", 30: "q_full 0", 31: "q_full 1", 32: "kv_full", 34: "S issued 0",
         35: "S issued 1", 36: "p_full 0", 37: "p_full 1", 38: "PV issued 0", 39: "PV issued 1"}
ROLE = ["softmax0", "softmax1", "loader", "issuer"]
buf = (C.c_ulonglong * (4 * 1024))()
cnt = (C.c_int * 4)()
cases = (("spatial", lambda: osb.attn_tiles(sp_t, sp_t, out, Lk=S, num_seqs=B * T)),
         ("temporal", lambda: osb.attn_tiles(tm_t, tm_t, out, Lk=T, num_seqs=B * S)),
         ("cross", lambda: osb.attn_tiles(q_t, kv_t, out, q_kind=0, k_kind=0, v_kind=1, Lk=Ly, num_seqs=B, kv_lens=lens)))
for name, fn in cases:
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    lib.osb_debug_ta_trace(buf, cnt)   # drop the warm-up launches' trace
    fn()
    torch.cuda.synchronize()
    assert lib.osb_debug_ta_trace(buf, cnt) == 0
    ev = []
    for r in range(4):
        for i in range(cnt[r]):
            ev.append((buf[r * 1024 + 2 * i + 1], r, buf[r * 1024 + 2 * i]))
    ev.sort()
    t0 = ev[0][0]
    print(f"=== {name}: {len(ev)} events, span {(ev[-1][0] - t0) / 1.85e3:.1f} us (clock64 / 1.85 GHz)")
    for t, r, tag in ev[:int(sys.argv[1]) if len(sys.argv) > 1 else 110]:
        print(f"{(t - t0) / 1.85e3:9.2f} us  {ROLE[r]:9s} {NAMES.get(int(tag), tag)}")
