"""Timeline of CTA 0 of the ping-pong attention kernel (instrumented build, tools/build_trace.sh): prints, per
STDiT3-XL/2 attention shape, the (role, tag, microseconds) events of the first jobs.  Debug aid, not a test."""
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
qkv = torch.randn(B * N, 3 * Cc, device="cuda").bfloat16()
qc = torch.randn(B * N, Cc, device="cuda").bfloat16()
kv = torch.randn(B * Ly, 2 * Cc, device="cuda").bfloat16()
out = torch.empty(B * N, Cc, device="cuda", dtype=torch.bfloat16)
w = torch.ones(D, device="cuda").bfloat16()
inv = 1.0 / (10000 ** (torch.arange(0, D, 2, device="cuda").float() / D))
ang = torch.arange(T, device="cuda").float()[:, None] * inv[None]
cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
lens = torch.tensor([260], device="cuda", dtype=torch.int32)


def spatial():
    osb.attn_short(qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:], out, num_seqs=B * T, seqs_per_batch=T, q_strides=(N, S, 1),
                   k_strides=(N, S, 1), Lq=S, Lk=S, num_heads=H, head_dim=D, q_norm_w=w, k_norm_w=w, impl=4)


def temporal():
    osb.attn_short(qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:], out, num_seqs=B * S, seqs_per_batch=S, q_strides=(N, 1, S),
                   k_strides=(N, 1, S), Lq=T, Lk=T, num_heads=H, head_dim=D, q_norm_w=w, k_norm_w=w, rope_cos=cos, rope_sin=sin, impl=4)


def cross():
    osb.attn_short(qc, kv[:, :Cc], kv[:, Cc:], out, num_seqs=B, seqs_per_batch=1, q_strides=(N, 0, 1), k_strides=(Ly, 0, 1),
                   Lq=N, Lk=Ly, num_heads=H, head_dim=D, kv_lens=lens, impl=4)


NAMES = {10: "s_full", 11: "max done", 12: "rescaled", 13: "P stored", 14: "o_full(final)", 15: "epilogue done",
         20: "first Q issued", 23: "Q copies landed", 24: "Q finished", 31: "KV copies landed", 32: "KV finished", 40: "q_full s0", 41: "q_full s1", 42: "kv_full s0", 43: "kv_full s1",
         44: "S issued s0", 45: "S issued s1", 46: "p_full s0", 47: "p_full s1", 48: "PV issued s0", 49: "PV issued s1"}
ROLE = ["softmax0", "softmax1", "loaders", "issuer"]
buf = (C.c_ulonglong * (4 * 512))()
cnt = (C.c_int * 4)()
for name, fn in (("spatial", spatial), ("temporal", temporal), ("cross", cross)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    lib.osb_debug_pp_trace(buf, cnt)   # drop the warm-up launches' trace
    fn()
    torch.cuda.synchronize()
    assert lib.osb_debug_pp_trace(buf, cnt) == 0
    ev = []
    for r in range(4):
        for i in range(cnt[r]):
            ev.append((buf[r * 512 + 2 * i + 1], r, buf[r * 512 + 2 * i]))
    ev.sort()
    t0 = ev[0][0]
    print(f"=== {name}: {len(ev)} events, span {(ev[-1][0] - t0) / 1.85e3:.1f} us (clock64 / 1.85 GHz)")
    for t, r, tag in ev[:int(sys.argv[1]) if len(sys.argv) > 1 else 140]:
        print(f"{(t - t0) / 1.85e3:9.2f} us  {ROLE[r]:9s} {NAMES.get(int(tag), tag)}")
