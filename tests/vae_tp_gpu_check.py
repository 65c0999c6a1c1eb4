"""Multi-GPU check (run under torchrun, NCCL): the frame-sharded VAE decode (`AutoencoderKLCausal3D.enable_temporal_parallel`:
two-frame causal halo from the left neighbour, group-wide GroupNorm statistics, one gather at the end) against the same model
decoding the whole latent on one GPU, on the real kernels, with the device time of both.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29513 tests/vae_tp_gpu_check.py

NOT yet run on a GPU box (written after the round's GPU budget was spent): the CPU twin on gloo ranks through the stand-in of
the binding is tests/test_host_vae_cpu.py::test_frame_sharded_decode_matches_the_whole_decode."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "open-sora_b200"))
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from opensora.registry import MODELS, build_module
    from tests.util import rel_l2

    ok = True
    for name, chans, groups, (T, H, W) in (("small", (32, 64, 64, 64), 8, (4 * world + 1, 16, 16)),
                                           ("hunyuan", (128, 256, 512, 512), 32, (max(9, 2 * world + 1), 32, 32))):
        torch.manual_seed(5)
        m = build_module(dict(type="hunyuan_vae", block_out_channels=chans, layers_per_block=1 if name == "small" else 2,
                              norm_num_groups=groups, latent_channels=16), MODELS, device_map="cpu")
        for p in m.parameters():      # the same model on every rank
            torch.nn.init.normal_(p, std=0.05) if p.dim() > 1 else None
        m = m.to("cuda", torch.bfloat16).eval()
        z = torch.randn(1, 16, T, H, W, generator=torch.Generator().manual_seed(7)).cuda()

        def timed(fn, reps=3):
            fn()
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                out = fn()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return out, t.item()

        with torch.no_grad():
            whole, t_whole = timed(lambda: m.decode(z))
            video = whole.clamp(-1, 1)
            lat, t_enc = timed(lambda: m.encode(video, sample_posterior=False))
            m.enable_temporal_parallel(dist.group.WORLD)
            sharded, t_shard = timed(lambda: m.decode(z))
            lat_sh, t_enc_sh = timed(lambda: m.encode(video, sample_posterior=False))
            m.enable_temporal_parallel(None)
        r = rel_l2(sharded, whole)
        re = rel_l2(lat_sh, lat)
        print(f"[vae-tp{world}] rank {rank} {name} encode {tuple(video.shape[2:])} -> {tuple(lat.shape[2:])}: rel_l2 vs whole encode = "
              f"{re:.3e}  whole {t_enc:.1f} ms  sharded {t_enc_sh:.1f} ms (x{t_enc / t_enc_sh:.2f})", flush=True)
        ok &= lat_sh.shape == lat.shape and re < 2e-2
        frames = whole.shape[2]
        print(f"[vae-tp{world}] rank {rank} {name} latent {T}x{H}x{W} -> {frames} frames: rel_l2 vs whole decode = {r:.3e}  "
              f"whole {t_whole:.1f} ms ({frames / t_whole * 1e3:.1f} fps)  sharded {t_shard:.1f} ms "
              f"({frames / t_shard * 1e3:.1f} fps, x{t_whole / t_shard:.2f})", flush=True)
        ok &= sharded.shape == whole.shape and r < 2e-2
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if t.item() != 1.0:
        sys.exit(1)
    if rank == 0:
        print("VAE_TP_CHECK_OK")


if __name__ == "__main__":
    main()
