"""Drop-in for `opensora/registry.py:1-41` without the mmengine dependency (absent in this image).

`build_module(cfg_dict, MODELS, **kwargs)` keeps the reference semantics (registry.py:7-30): deep-copy
the dict, inject kwargs, pop `type`, call the registered callable; an nn.Module is passed through."""
from __future__ import annotations

import importlib
from copy import deepcopy

import torch.nn as nn


class Registry:
    """Minimal mmengine.registry.Registry: register_module(name) decorator, get(), build(cfg)."""

    def __init__(self, name: str, locations: list[str] | None = None):
        self.name = name
        self.locations = locations or []
        self._module_dict: dict[str, object] = {}
        self._imported = False

    def _import_locations(self) -> None:
        if self._imported:
            return
        self._imported = True
        for loc in self.locations:
            try:
                importlib.import_module(loc)
            except ModuleNotFoundError as e:  # e.g. opensora.datasets is out of scope
                if e.name != loc:
                    raise

    def register_module(self, name: str | list[str] | None = None, force: bool = False, module=None):
        def _register(obj):
            names = [name] if isinstance(name, str) else (name or [obj.__name__])
            for n in names:
                if not force and n in self._module_dict:
                    raise KeyError(f"{n} is already registered in {self.name}")
                self._module_dict[n] = obj
            return obj

        if module is not None:
            return _register(module)
        return _register

    def get(self, key: str):
        self._import_locations()
        return self._module_dict.get(key)

    def build(self, cfg: dict):
        cfg = dict(cfg)
        if "type" not in cfg:
            raise KeyError(f'`cfg` must contain the key "type", but got {cfg}')
        obj_type = cfg.pop("type")
        obj = self.get(obj_type) if isinstance(obj_type, str) else obj_type
        if obj is None:
            raise KeyError(f"{obj_type} is not in the {self.name} registry")
        return obj(**cfg)

    def __contains__(self, key: str) -> bool:
        return self.get(key) is not None


def build_module(module: dict | nn.Module | None, builder: Registry, **kwargs) -> nn.Module | None:
    if module is None:
        return None
    if isinstance(module, dict):
        cfg = deepcopy(module)
        for k, v in kwargs.items():
            cfg[k] = v
        return builder.build(cfg)
    if isinstance(module, nn.Module):
        return module
    raise TypeError(f"Only support dict and nn.Module, but got {type(module)}.")


MODELS = Registry("model", locations=["opensora.models"])
DATASETS = Registry("dataset", locations=["opensora.datasets"])
SCHEDULERS = Registry("scheduler", locations=["opensora.schedulers"])   # v1.2 name; the v2.0 tree has no scheduler registry
