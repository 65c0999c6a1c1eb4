"""Inference-only shim of `opensora/acceleration/checkpoint.py:254-271`: activation checkpointing is a
training feature (out of scope); the call sites only need the pass-through."""


def set_grad_checkpoint(model, use_fp32_attention=False, gc_step=1):
    raise NotImplementedError("osb200 is a forward-only (inference) path; activation checkpointing is out of scope")


def auto_grad_checkpoint(module, *args, **kwargs):
    return module(*args, **kwargs)
