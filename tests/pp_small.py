"""Small / ragged shapes of the XS parity test through the ping-pong attention kernel, one per process
(usage: pp_small.py spatial|temporal|cross [impl])."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-sora_b200"))
sys.path.insert(0, ROOT)
import torch

import osb200 as osb
from tests.test_kernels_gpu import _attn_ref, _randn, _rope_tables

osb.init(0)
mode = sys.argv[1]
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
B, T, S, H, D, Ly = 2, 4, 24, 4, 72, 300
C, N = H * D, T * S
qkv = _randn(B * N, 3 * C, seed=41)
qw, kw = _randn(D, seed=42) * 0.2 + 1, _randn(D, seed=43) * 0.2 + 1
out = torch.zeros(B * N, C, dtype=torch.bfloat16, device="cuda")
q2, k2, v2 = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
if mode == "spatial":
    st = (N, S, 1)
    osb.attn_short(q2, k2, v2, out, num_seqs=B * T, seqs_per_batch=T, q_strides=st, k_strides=st, Lq=S, Lk=S,
                   num_heads=H, head_dim=D, q_norm_w=qw, k_norm_w=kw, impl=impl)
    x = qkv.float().view(B * T, S, 3, H, D).permute(2, 0, 3, 1, 4)
    ref = _attn_ref(x[0], x[1], x[2], qw, kw, None, None, D ** -0.5).permute(0, 2, 1, 3).reshape(B * N, C)
elif mode == "temporal":
    cos, sin = _rope_tables(T, D)
    st = (N, 1, S)
    osb.attn_short(q2, k2, v2, out, num_seqs=B * S, seqs_per_batch=S, q_strides=st, k_strides=st, Lq=T, Lk=T,
                   num_heads=H, head_dim=D, q_norm_w=qw, k_norm_w=kw, rope_cos=cos, rope_sin=sin, impl=impl)
    x = qkv.float().view(B, T, S, 3, H, D).permute(3, 0, 2, 4, 1, 5).reshape(3, B * S, H, T, D)
    ref = _attn_ref(x[0], x[1], x[2], qw, kw, cos, sin, D ** -0.5)
    ref = ref.view(B, S, H, T, D).permute(0, 3, 1, 2, 4).reshape(B * N, C)
else:
    q = _randn(B * N, C, seed=44)
    kv = _randn(B * Ly, 2 * C, seed=45)
    lens = torch.tensor([300, 289], device="cuda", dtype=torch.int32)
    osb.attn_short(q, kv[:, :C], kv[:, C:], out, num_seqs=B, seqs_per_batch=1, q_strides=(N, 0, 1), k_strides=(Ly, 0, 1),
                   Lq=N, Lk=Ly, num_heads=H, head_dim=D, kv_lens=lens, impl=impl)
    qf = q.float().view(B, N, H, D).permute(0, 2, 1, 3)
    kvf = kv.float().view(B, Ly, 2, H, D).permute(2, 0, 3, 1, 4)
    ref = _attn_ref(qf, kvf[0], kvf[1], None, None, None, None, D ** -0.5, kv_len=lens).permute(0, 2, 1, 3).reshape(B * N, C)
torch.cuda.synchronize()
r = float((out.float() - ref).norm() / ref.norm())
print(f"pp_small {mode} impl{impl}: rel_l2 {r:.3e}")
