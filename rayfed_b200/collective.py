"""``fed.aggregate`` — cross-party mean of per-party values as ONE DAG-level call.

The reference has no such primitive: averaging is user code that gathers every
party's value to one party, takes the mean and sends it back
(fed/tests/test_fed_get.py:45-47,71-78) — 2(N-1) pickled copies per round.  Here::

    new = fed.aggregate([w_alice, w_bob, w_carol])      # FedObjects, one per party
    for p, w in zip(parties, new):
        trainers[p].set_weights.remote(w)                # w is owned by party p: no transfer

* **fused path** (every party has a device channel — ``fed.init(config={"device": ...})``):
  each party's value stays on its GPU; the party's hidden aggregation actor stages the
  tensor leaves in a symmetric bf16 bucket and launches the fused aggregate kernel
  (reduce over peers + weights + all-gather inside one kernel, K2).  No value crosses the
  DAG at all — the returned FedObjects are owned by the same parties as the inputs.
  Precision: fp32 accumulation, bf16 on the wire and in the result (cast back to each
  leaf's dtype), exactly like :class:`rayfed_b200.fl.FedAvgWorker`.
* **DAG path** (no device, or ``fused=False``): gather to the first party, weighted mean in
  the leaves' own dtype, redistribute — the reference's pattern, for CPU jobs and for
  values that must not be rounded to bf16.

Like every fed call this is executed by all controllers; it consumes the same number of
sequence ids everywhere (the hidden actors are created on first use in every party).
"""
from __future__ import annotations

import threading
from typing import Dict, List, Optional, Sequence

from .fed_object import FedObject
from .tree_util import tree_flatten, tree_unflatten

_STATE_LOCK = threading.Lock()
_ACTORS: Dict[str, Dict[str, object]] = {}  # job -> {party: FedActorHandle}


def _weighted_mean(weights, *values):
    """Leaf-wise weighted mean of equally-structured pytrees (host/DAG path)."""
    flats, spec0 = [], None
    for v in values:
        leaves, spec = tree_flatten(v)
        if spec0 is None:
            spec0 = spec
        elif len(leaves) != len(flats[0]):
            raise ValueError("fed.aggregate: the parties' values do not have the same structure")
        flats.append(leaves)
    n = len(values)
    w = list(weights) if weights is not None else [1.0 / n] * n
    out = []
    for column in zip(*flats):
        first = column[0]
        if hasattr(first, "dtype") and hasattr(first, "shape") and not isinstance(first, (int, float)):
            try:
                import torch
            except ImportError:  # pragma: no cover - numpy-only environments
                torch = None
            if torch is not None and isinstance(first, torch.Tensor):
                acc = torch.zeros_like(first, dtype=torch.float32 if first.is_floating_point() else torch.float64)
                for wi, x in zip(w, column):
                    acc += wi * x.to(device=first.device, dtype=acc.dtype)
                out.append(acc.to(first.dtype) if first.is_floating_point() else acc)
            else:  # numpy and friends
                acc = sum(wi * x for wi, x in zip(w, column))
                out.append(acc)
        else:
            out.append(sum(wi * x for wi, x in zip(w, column)))
    return tree_unflatten(out, spec0)


def _identity(x):
    return x


class _PartyAggregator:
    """Hidden per-party actor of the fused path: owns the staging buckets (one per value
    signature) and serialises this party's aggregate launches in program order — the same
    order in every party, which is what makes the in-kernel rendezvous safe."""

    def __init__(self, parties: Sequence[str], party: str):
        from .fl.fedavg import device_context

        self.parties = list(parties)
        self.party = party
        self.group, self.heap, self.channel = device_context(self.parties, party)
        self._slots = {}

    def _slot(self, leaves, weights):
        import torch

        from .fl.bucket import FlatBucket
        from .fl.fedavg import FedAggregator

        sig = (tuple((tuple(x.shape), str(x.dtype)) for x in leaves), tuple(weights) if weights else None)
        slot = self._slots.get(sig)
        if slot is None:
            dev = torch.device(f"cuda:{self.group.device_index}")
            holder = torch.nn.Module()
            holder.leaves = torch.nn.ParameterList(
                [torch.nn.Parameter(torch.zeros(x.shape, dtype=torch.bfloat16, device=dev)) for x in leaves])
            name = f"fed-aggregate/{len(self._slots)}"
            bucket = FlatBucket(holder, self.heap, name=name, n_parties=self.group.size)
            agg = FedAggregator(bucket, self.group, self.channel, kind="fedavg",
                                weights=list(weights) if weights else None, heap=self.heap)
            slot = self._slots[sig] = (holder, bucket, agg)
        return slot

    def mean(self, value, weights=None):
        import torch

        leaves, spec = tree_flatten(value)
        dev = torch.device(f"cuda:{self.group.device_index}")
        tensors = []
        for x in leaves:
            t = x if isinstance(x, torch.Tensor) else torch.as_tensor(x)
            if not t.is_floating_point():
                t = t.to(torch.float32)
            tensors.append(t.to(dev))
        holder, bucket, agg = self._slot(tensors, weights)
        with torch.no_grad():
            for p, t in zip(holder.leaves, tensors):
                p.data.copy_(t)  # cast to bf16 into the symmetric bucket
        agg.aggregate()  # collective: all-reduce + weights + all-gather in one kernel
        if agg.status() != 0:
            from .exceptions import FedError

            raise FedError("fed.aggregate: a peer did not reach the in-kernel barrier in time")
        out = []
        for p, t, x in zip(holder.leaves, tensors, leaves):
            y = p.data.to(t.dtype).clone()
            if not isinstance(x, torch.Tensor):
                y = y.cpu()
                y = y.item() if y.dim() == 0 else y.numpy()
            elif not x.is_cuda:
                y = y.to(x.device)
            out.append(y)
        return tree_unflatten(out, spec)


def _job_actors(fed_api, parties, job_name):
    with _STATE_LOCK:
        actors = _ACTORS.get(job_name)
        if actors is None:
            cls = fed_api.remote(_PartyAggregator)
            actors = {p: cls.party(p).remote(parties, p) for p in parties}
            _ACTORS[job_name] = actors
        return actors


def reset(job_name: Optional[str] = None) -> None:
    """Forget the hidden actors (called by ``fed.shutdown``)."""
    with _STATE_LOCK:
        if job_name is None:
            _ACTORS.clear()
        else:
            _ACTORS.pop(job_name, None)


def aggregate(fed_objects: Sequence[FedObject], weights: Optional[Sequence[float]] = None,
              fused: Optional[bool] = None) -> List[FedObject]:
    """Weighted mean across parties.  ``fed_objects``: exactly one FedObject per party of the
    job (any order; pytrees of tensors / arrays / numbers with identical structure).
    ``weights``: per-object weights in the same order (default ``1/N`` each; they are used as
    given, normalise them yourself for sample-count weighting).  Returns one FedObject per
    input, owned by the same party, holding the mean.  See the module docstring for
    ``fused``."""
    import sys

    from . import api as fed_api
    from ._private.global_context import get_global_context

    ctx = get_global_context()
    if ctx is None:
        raise RuntimeError("fed.aggregate needs fed.init() first")
    job = ctx.get_job_name()
    parties = list(fed_api._get_addresses(job))
    objs = list(fed_objects)
    owners = [o.get_party() for o in objs]
    if sorted(owners) != sorted(parties):
        raise ValueError(f"fed.aggregate needs exactly one FedObject per party {parties}, got owners {owners}")
    if weights is not None and len(weights) != len(objs):
        raise ValueError("fed.aggregate: len(weights) must equal len(fed_objects)")

    nv = sys.modules.get("rayfed_b200.proxy.nvlink.nvlink_proxy")  # imported iff the NVLink transport is in use
    have_channel = nv is not None and nv.current_channel() is not None
    if fused is None:
        fused = have_channel
    if fused:
        if not have_channel:
            raise RuntimeError("fed.aggregate(fused=True) needs the NVLink device channel "
                               "(fed.init(config={'device': 'cuda:N', ...}) with >= 2 parties)")
        actors = _job_actors(fed_api, parties, job)
        # the kernel indexes weights by party rank, not by argument position
        by_rank = None
        if weights is not None:
            w_of = dict(zip(owners, weights))
            by_rank = [float(w_of[p]) for p in parties]
        return [actors[o.get_party()].mean.remote(o, by_rank) for o in objs]

    server = parties[0]
    mean = fed_api.remote(_weighted_mean).party(server).remote(list(weights) if weights is not None else None, *objs)
    ident = fed_api.remote(_identity)
    return [mean if o.get_party() == server else ident.party(o.get_party()).remote(mean) for o in objs]
