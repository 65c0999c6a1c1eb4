"""world_size-2 gloo tests (CPU) of the sequence-parallel collectives: the T-shard <-> S-shard
transposition used around every temporal attention, and the entry split / exit gather."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from opensora.acceleration.communications import (all_to_all, gather_forward_split_backward,
                                                          split_forward_gather_backward)

        g = dist.group.WORLD
        B, T, S, C = 2, 4, 6, 5
        full = torch.arange(B * T * S * C, dtype=torch.float32).view(B, T, S, C)
        # entry: T-sharded tokens
        loc = split_forward_gather_backward(full, g, dim=1)
        assert torch.equal(loc, full[:, rank * T // world:(rank + 1) * T // world])
        # transposition to S-sharded: scatter S, gather T
        s_sh = all_to_all(loc, g, scatter_dim=2, gather_dim=1)
        assert torch.equal(s_sh, full[:, :, rank * S // world:(rank + 1) * S // world])
        # and back
        back = all_to_all(s_sh, g, scatter_dim=1, gather_dim=2)
        assert torch.equal(back, loc)
        # exit gather
        out = gather_forward_split_backward(loc, g, dim=1)
        assert torch.equal(out, full)
        # flattened-token form used by the model: rows [B*Tl*S, C]
        rows = loc.reshape(B * (T // world) * S, C)
        s_rows = all_to_all(rows.view(B, T // world, S, C), g, 2, 1).reshape(-1, C)
        assert torch.equal(s_rows.view(B, T, S // world, C), full[:, :, rank * S // world:(rank + 1) * S // world])
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sp_collectives_world2():
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)


def test_single_rank_is_identity():
    from opensora.acceleration.communications import all_to_all, gather_forward_split_backward

    x = torch.randn(2, 3, 4)
    assert all_to_all(x, None) is x
    assert gather_forward_split_backward(x, None, 1) is x
