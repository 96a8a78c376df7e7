"""Round-state checkpoint / resume for federated jobs.

The reference persists nothing (the seq counter restarts at 0 on
re-init and configs live only in Ray's KV).  A federated *job* on this framework
can checkpoint per party: the flat weight bucket, the sharded optimizer state of
the aggregator, the local optimizer state and the round counter.  Files are
written atomically (tmp + rename); ``latest`` resolves the newest complete one.
"""
from __future__ import annotations

import glob
import os
from typing import Optional

import torch


def save_round_state(directory: str, party: str, round_idx: int, *, bucket=None, aggregator=None, local_opt=None,
                     extra: Optional[dict] = None) -> str:
    os.makedirs(directory, exist_ok=True)
    state = {"party": party, "round": round_idx, "extra": extra or {}}
    if bucket is not None:
        state["flat_w"] = bucket.flat_w.detach().cpu()
    if aggregator is not None:
        state["agg"] = {
            "step_count": aggregator.step_count,
            "master": None if aggregator.master is None else aggregator.master.cpu(),
            "mom": None if aggregator.mom is None else aggregator.mom.cpu(),
            "var": None if aggregator.var is None else aggregator.var.cpu(),
        }
    if local_opt is not None:
        state["local_mom"] = local_opt.mom.cpu()
    path = os.path.join(directory, f"{party}.round{round_idx:08d}.pt")
    tmp = path + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, path)
    return path


def latest(directory: str, party: str) -> Optional[str]:
    files = sorted(glob.glob(os.path.join(directory, f"{party}.round*.pt")))
    return files[-1] if files else None


def load_round_state(path: str, *, bucket=None, aggregator=None, local_opt=None) -> dict:
    state = torch.load(path, map_location="cpu", weights_only=False)
    if bucket is not None and "flat_w" in state:
        bucket.flat_w.copy_(state["flat_w"].to(bucket.flat_w.device))
    if aggregator is not None and "agg" in state:
        a = state["agg"]
        aggregator.step_count = a["step_count"]
        for name in ("master", "mom", "var"):
            t = getattr(aggregator, name)
            if t is not None and a[name] is not None:
                t.copy_(a[name].to(t.device))
    if local_opt is not None and "local_mom" in state:
        local_opt.mom.copy_(state["local_mom"].to(local_opt.mom.device))
    return state
