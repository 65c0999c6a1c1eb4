"""Forward-only sequence-parallel collectives with the reference's public names and argument meaning
(`opensora/acceleration/communications.py:57-63` all_to_all, `:183-188` split/gather wrappers) — the
scheme north_star calls "the repo's own".  The reference versions are autograd Functions over
list-based `dist.all_to_all` / `all_gather` (`:8-18`, `:83-120`); inference needs only the forward, done
here with ONE `all_to_all_single` / `all_gather_into_tensor` on a layout where the exchange is a plain
leading-dim split (one packing copy in, one un-packing copy out when the dims are not leading)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def _world(group) -> int:
    return dist.get_world_size(group) if group is not None else 1


def all_to_all(input_: torch.Tensor, process_group, scatter_dim: int = 2, gather_dim: int = 1) -> torch.Tensor:
    """Split `input_` in P chunks along `scatter_dim`, exchange, concatenate the received chunks along
    `gather_dim` (semantics of communications.py:8-18)."""
    P = _world(process_group)
    if P == 1:
        return input_
    scatter_dim %= input_.dim()
    gather_dim %= input_.dim()
    n = input_.size(scatter_dim)
    assert n % P == 0, f"scatter dim {scatter_dim} of size {n} is not divisible by world size {P}"
    # [.., P*c, ..] -> [P, .., c, ..] contiguous: chunk p goes to rank p
    shp = list(input_.shape)
    shp[scatter_dim:scatter_dim + 1] = [P, n // P]
    send = input_.reshape(shp).movedim(scatter_dim, 0).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=process_group)
    # recv[p] is the chunk from rank p: concatenate along gather_dim
    out = recv.movedim(0, gather_dim)  # [.., P, g, ..]
    shp = list(out.shape)
    shp[gather_dim:gather_dim + 2] = [shp[gather_dim] * shp[gather_dim + 1]]
    return out.reshape(shp).contiguous()


def _split(input_: torch.Tensor, pg, dim: int = -1) -> torch.Tensor:
    P = _world(pg)
    if P == 1:
        return input_
    n = input_.size(dim)
    assert n % P == 0, (f"The dimension to split ({n}) is not a multiple of world size ({P}), "
                        f"cannot split tensor evenly")
    return torch.split(input_, n // P, dim=dim)[dist.get_rank(pg)].contiguous()


def _gather(input_: torch.Tensor, pg, dim: int = -1) -> torch.Tensor:
    P = _world(pg)
    if P == 1:
        return input_
    input_ = input_.contiguous()
    flat = torch.empty((P * input_.shape[0],) + tuple(input_.shape[1:]), dtype=input_.dtype, device=input_.device)
    dist.all_gather_into_tensor(flat, input_, group=pg)  # rank-major concatenation along dim 0
    out = flat.view((P,) + tuple(input_.shape))
    dim %= input_.dim()
    out = out.movedim(0, dim)
    shp = list(out.shape)
    shp[dim:dim + 2] = [shp[dim] * shp[dim + 1]]
    return out.reshape(shp).contiguous()


def split_forward_gather_backward(input_, process_group, dim, grad_scale=1.0):
    """communications.py:183-185 — forward = keep this rank's chunk (no communication)."""
    return _split(input_, process_group, dim)


def gather_forward_split_backward(input_, process_group, dim, grad_scale=None):
    """communications.py:187-188 — forward = all-gather and concatenate along `dim`."""
    return _gather(input_, process_group, dim)


def gather_forward_split_backward_var_len(input_: torch.Tensor, dim: int, process_group, splits) -> torch.Tensor:
    """Forward of the reference's var-len gather (`opensora/models/mmdit/distributed.py:671-679`): rank r holds
    `splits[r]` entries along `dim`; every rank ends up with the concatenation in rank order.  Chunks are padded to the
    longest one for ONE `all_gather_into_tensor`."""
    P = _world(process_group)
    if P == 1:
        return input_
    dim %= input_.dim()
    mx = max(int(v) for v in splits)
    x = input_.movedim(dim, 0).contiguous()
    assert x.shape[0] == int(splits[dist.get_rank(process_group)])
    if x.shape[0] < mx:
        x = torch.cat([x, x.new_zeros((mx - x.shape[0],) + tuple(x.shape[1:]))], 0)
    flat = torch.empty((P * mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(flat, x, group=process_group)
    parts = [flat[r * mx:r * mx + int(splits[r])] for r in range(P)]
    return torch.cat(parts, 0).movedim(0, dim).contiguous()
