"""FedCallHolder — submits one DAG node (parity: fed/_private/fed_call_holder.py:62-110).

Every controller executes the same call; the holder mints the node's sequence
id and then takes one of two branches:

* **owner** (this party runs the task): resolve FedObject arguments to local
  futures (receiving remote ones), submit to the party runtime, wrap the result
  future(s) in FedObject(s);
* **non-owner**: push every argument *owned here* to the owner — at most once
  per (object, destination) — and return placeholder FedObject(s).
"""
from __future__ import annotations

import logging

from ..fed_object import FedObject
from ..proxy.barriers import send
from ..tree_util import has_containers, tree_flatten
from ..utils import resolve_dependencies
from .global_context import get_global_context

logger = logging.getLogger(__name__)


class FedCallHolder:
    """Binds (party, submit function, options) for one ``.remote()`` call."""

    __slots__ = ("_party", "_node_party", "_options", "_submit_ray_task_func")

    def __init__(self, node_party, submit_ray_task_func, options=None) -> None:
        ctx = get_global_context()
        # current party is a per-job constant: read it from the context, not the KV
        self._party = ctx.get_current_party() if ctx is not None else None
        self._node_party = node_party
        self._options = options or {}
        self._submit_ray_task_func = submit_ray_task_func

    def options(self, **options):
        self._options = options
        return self

    def internal_remote(self, *args, **kwargs):
        if not self._node_party:
            raise ValueError("You should specify a party name on the fed actor.")
        ctx = get_global_context()
        fed_task_id = ctx.next_seq_id()
        num_returns = self._options.get("num_returns", 1) if self._options else 1

        if self._party == self._node_party:
            resolved_args, resolved_kwargs = resolve_dependencies(self._party, fed_task_id, *args, **kwargs)
            refs = self._submit_ray_task_func(resolved_args, resolved_kwargs)
            if isinstance(refs, list):
                return [FedObject(self._node_party, fed_task_id, ref, i) for i, ref in enumerate(refs)]
            return FedObject(self._node_party, fed_task_id, refs)

        # Not ours to run: feed the owner whatever it needs from us.
        if has_containers(args, kwargs):
            leaves, _ = tree_flatten((args, kwargs))
        else:
            leaves = list(args) + (list(kwargs.values()) if kwargs else [])
        for arg in leaves:
            if type(arg) is FedObject and arg.get_party() == self._party:
                if arg._was_sending_or_sent_to_party(self._node_party):
                    continue
                arg._mark_is_sending_to_party(self._node_party)
                send(
                    dest_party=self._node_party,
                    data=arg.get_ref(),
                    upstream_seq_id=arg.get_fed_task_id(),
                    downstream_seq_id=fed_task_id,
                )
        if num_returns and num_returns > 1:
            return [FedObject(self._node_party, fed_task_id, None, i) for i in range(num_returns)]
        return FedObject(self._node_party, fed_task_id, None)
