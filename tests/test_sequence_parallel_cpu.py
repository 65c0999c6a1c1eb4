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


@pytest.mark.parametrize("P,B,T,S", [(2, 2, 4, 6), (4, 1, 8, 12), (8, 1, 64, 256)])
def test_peer_scatter_routing_equals_all_to_all(P, B, T, S):
    """The row routing the kernels use for the peer-memory exchange (osb_scatter) reproduces the reference's all_to_all
    semantics (communications.py:8-18): simulated for P ranks in one process, both directions."""
    from opensora.acceleration.peer_exchange import scatter_dest

    C, Tl, Sl = 3, T // P, S // P
    full = torch.arange(B * T * S * C, dtype=torch.float32).view(B, T, S, C)
    # T-sharded -> S-sharded (mode 1): producers hold [B, Tl, S], consumers must end up with full[:, :, p*Sl:(p+1)*Sl]
    recv = [torch.full((B * T * Sl, C), -1.0) for _ in range(P)]
    for r in range(P):
        src = full[:, r * Tl:(r + 1) * Tl].reshape(B * Tl * S, C)
        for row in range(src.shape[0]):
            p, d = scatter_dest(1, P, r, Tl, S, row)
            recv[p][d] = src[row]
    for p in range(P):
        assert torch.equal(recv[p].view(B, T, Sl, C), full[:, :, p * Sl:(p + 1) * Sl])
    # S-sharded -> T-sharded (mode 2): producers hold [B, T, Sl], consumers must end up with full[:, p*Tl:(p+1)*Tl]
    back = [torch.full((B * Tl * S, C), -1.0) for _ in range(P)]
    for r in range(P):
        src = full[:, :, r * Sl:(r + 1) * Sl].reshape(B * T * Sl, C)
        for row in range(0, src.shape[0], 1 if src.shape[0] < 4096 else 7):   # sampled at the large shape
            p, d = scatter_dest(2, P, r, T, Sl, row)
            back[p][d] = src[row]
    for p in range(P):
        want = full[:, p * Tl:(p + 1) * Tl].reshape(B * Tl * S, C)
        hit = back[p][:, 0] >= 0
        assert hit.any() and torch.equal(back[p][hit], want[hit])


def test_transposing_scatter_modes():
    """Host mirror of osb_scatter modes 3 (local transpose) and 4 (exchange with transposed destination)."""
    from opensora.acceleration.peer_exchange import scatter_dest

    P, B, T, S = 4, 2, 8, 12
    Tl, Sl = T // P, S // P
    full = torch.arange(B * T * S, dtype=torch.float32).view(B, T, S)
    t3 = torch.full((B * S * T,), -1.0)
    for row in range(B * T * S):
        p, d = scatter_dest(3, 1, 0, T, S, row)
        assert p == 0
        t3[d] = full.reshape(-1)[row]
    assert torch.equal(t3.view(B, S, T), full.transpose(1, 2))
    recv = [torch.full((B * Sl * T,), -1.0) for _ in range(P)]
    for r in range(P):
        src = full[:, r * Tl:(r + 1) * Tl].reshape(-1)
        for row in range(src.numel()):
            p, d = scatter_dest(4, P, r, Tl, S, row)
            recv[p][d] = src[row]
    for p in range(P):
        assert torch.equal(recv[p].view(B, Sl, T), full[:, :, p * Sl:(p + 1) * Sl].transpose(1, 2))
