"""Checkpoint files in the reference's format.

The reference writes `torch.save({"model": model.state_dict(), "optimizer": optimizer.state_dict(), "it": ..., "t": ...,
"run_id": ..., <best metrics>}, "<out_dir>/model.pt" | "best_model_all.pt")` (utils/checkpoint.py:31-47 called from
experiment.py:121-128, 242-254, 390) and reads it back module by module (utils/checkpoint.py:49-72).  Because the drop-in
`ViewFusion` / `UNet` keep the reference's 406 state_dict keys and `FusedAdam` keeps `torch.optim.Adam`'s state layout,
those files load here unchanged — these helpers only take care of the two things that differ in practice:
  * files saved from a `DistributedDataParallel` wrapper carry a "module." prefix on every key;
  * the packed bf16 weights the CUDA path reads are derived data: they are refreshed lazily because `load_state_dict`
    bumps the parameters' version counters (`UNet.packed_weights`), nothing has to be done by the caller.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch


def _strip_prefix(sd: Dict[str, Any], prefix: str = "module.") -> Dict[str, Any]:
    if sd and all(k.startswith(prefix) for k in sd):
        return {k[len(prefix):]: v for k, v in sd.items()}
    return sd


def save_checkpoint(path: str, model, optimizer=None, **extra) -> None:
    """Same file layout as utils/checkpoint.py:31-47 with `model=` and `optimizer=` registered."""
    out = dict(extra)
    out["model"] = getattr(model, "module", model).state_dict()
    if optimizer is not None:
        out["optimizer"] = optimizer.state_dict()
    torch.save(out, path)


def load_checkpoint(path: str, model, optimizer=None, map_location=None, strict: bool = True) -> Dict[str, Any]:
    """Loads `model` (and `optimizer` when given and present) from a reference-format file; returns the remaining entries
    (`it`, `t`, `run_id`, best metrics), like utils/checkpoint.py:49-72.  Raises KeyError if the file has no "model"."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    if "model" not in ckpt:
        raise KeyError(f'{path}: no "model" entry (keys: {sorted(ckpt)})')
    target = getattr(model, "module", model)
    target.load_state_dict(_strip_prefix(ckpt["model"]), strict=strict)
    if optimizer is not None and "optimizer" in ckpt:
        optimizer.load_state_dict(ckpt["optimizer"])
    return {k: v for k, v in ckpt.items() if k not in ("model", "optimizer")}
