"""First-contact diagnostics for the tcgen05 kernels (not a pytest file): prints small slices of
kernel output next to the fp32 reference so a descriptor / swizzle mistake can be localised from
one gpurun round trip.  Usage: python tests/gpu_debug.py <case> [...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "open-sora_b200"))
import torch

import osb200 as osb

torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)


def stats(name, out, ref):
    out, ref = out.float(), ref.float()
    err = (out - ref).abs()
    rel = float((out - ref).norm() / ref.norm())
    print(f"[{name}] rel_l2={rel:.3e} max_abs={float(err.max()):.3e} nan={int(torch.isnan(out).sum())} "
          f"zeros={float((out == 0).float().mean()):.3f}")
    return rel


def gemm_case(M, N, K, cta, bn=0, pattern="rand"):
    torch.manual_seed(0)
    if pattern == "rand":
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    else:  # structured: A[m,k] = m, W[n,k] = (k==n) -> out[m,n] = A[m,n]
        a = (torch.arange(M, device="cuda")[:, None] % 64 + torch.arange(K, device="cuda")[None, :] * 0.0).bfloat16()
        a = a + (torch.arange(K, device="cuda")[None, :] % 8).bfloat16() * 64
        w = torch.zeros(N, K, device="cuda").bfloat16()
        idx = torch.arange(min(N, K), device="cuda")
        w[idx, idx] = 1
    out = osb.gemm(a, w, None, cta_group=cta, block_n=bn)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    rel = stats(f"gemm {M}x{N}x{K} cta{cta} bn{bn} {pattern}", out, ref)
    if rel > 2e-3:
        print("out[:8,:8]\n", out[:8, :8].float().cpu())
        print("ref[:8,:8]\n", ref[:8, :8].cpu())
        rowerr = (out.float() - ref).abs().amax(dim=1)
        colerr = (out.float() - ref).abs().amax(dim=0)
        print("bad rows:", torch.nonzero(rowerr > 0.05).flatten()[:40].tolist())
        print("bad cols:", torch.nonzero(colerr > 0.05).flatten()[:40].tolist())
    return rel


def main():
    osb.init(0)
    print("device", torch.cuda.get_device_name(0), "osb version", osb.version())
    case = sys.argv[1] if len(sys.argv) > 1 else "gemm1"
    if case == "ln":
        x = torch.randn(100, 1152, device="cuda").bfloat16()
        mod = torch.randn(1, 2, 1152, device="cuda")
        y = osb.ln_modulate(x, mod[:, 0], mod[:, 1], group_rows=100)
        ref = torch.nn.functional.layer_norm(x.float(), (1152,), eps=1e-6) * (1 + mod[0, 1]) + mod[0, 0]
        stats("ln_modulate", y, ref)
    elif case == "gemm1":
        gemm_case(128, 64, 64, 1, 64, "eye")
        gemm_case(128, 64, 64, 1, 64)
        gemm_case(128, 128, 256, 1, 128)
        gemm_case(256, 256, 512, 1, 256)
        gemm_case(1024, 1152, 1152, 1)
        gemm_case(16384, 3456, 1152, 1)
    elif case == "gemm2":
        gemm_case(256, 64, 64, 2, 64, "eye")
        gemm_case(256, 64, 64, 2, 64)
        gemm_case(256, 256, 512, 2, 256)
        gemm_case(1024, 1152, 1152, 2)
        gemm_case(16384, 3456, 1152, 2)
    elif case == "attn":
        D, H, S, nseq = 72, 2, 256, 3
        C = H * D
        torch.manual_seed(0)
        qkv = torch.randn(nseq * S, 3 * C, device="cuda").bfloat16()
        out = torch.zeros(nseq * S, C, dtype=torch.bfloat16, device="cuda")
        osb.attn_short(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, num_seqs=nseq, seqs_per_batch=nseq,
                       q_strides=(0, S, 1), k_strides=(0, S, 1), Lq=S, Lk=S, num_heads=H, head_dim=D)
        torch.cuda.synchronize()
        x = qkv.float().view(nseq, S, 3, H, D).permute(2, 0, 3, 1, 4)
        s = (x[0] @ x[1].transpose(-1, -2)) * D ** -0.5
        ref = (torch.softmax(s, -1) @ x[2]).permute(0, 2, 1, 3).reshape(nseq * S, C)
        rel = stats("attn S256 D72", out, ref)
        if rel > 4e-3:
            print("out[:4,:12]\n", out[:4, :12].float().cpu(), "\nref[:4,:12]\n", ref[:4, :12].cpu())
    print("launches", osb.launch_count())


if __name__ == "__main__":
    main()
