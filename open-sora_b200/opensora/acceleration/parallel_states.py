"""Process-group registry behind the accessor names the reference's models and scripts call
(`opensora/acceleration/parallel_states.py:6-29`): `set_/get_{data,sequence,tensor}_parallel_group`.

One table keyed by parallelism kind; the accessors are generated from it.  Defaults follow upstream: an unset data
group means the whole world (or the "mixed" data group when a hybrid plugin registered one and the caller asks for it),
an unset sequence / tensor group means "not parallel" (None)."""
from __future__ import annotations

import torch.distributed as dist


class _GroupTable:
    KINDS = ("data", "sequence", "tensor", "mixed_dp_group")

    def __init__(self):
        self._groups = {}

    def put(self, kind: str, group) -> None:
        if kind not in self.KINDS:
            raise KeyError(f"unknown parallel group kind {kind!r}")
        self._groups[kind] = group

    def find(self, kind: str, default=None):
        return self._groups[kind] if kind in self._groups else default

    def clear(self) -> None:
        self._groups.clear()


_TABLE = _GroupTable()


def _setter(kind: str):
    def set_group(group) -> None:
        _TABLE.put(kind, group)

    set_group.__name__ = f"set_{kind}_parallel_group"
    return set_group


set_data_parallel_group = _setter("data")
set_sequence_parallel_group = _setter("sequence")
set_tensor_parallel_group = _setter("tensor")


def get_data_parallel_group(get_mixed_dp_pg: bool = False):
    mixed = _TABLE.find("mixed_dp_group") if get_mixed_dp_pg else None
    return mixed if mixed is not None else _TABLE.find("data", dist.group.WORLD)


def get_sequence_parallel_group():
    return _TABLE.find("sequence")


def get_tensor_parallel_group():
    return _TABLE.find("tensor")
