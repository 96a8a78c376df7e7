"""FedObject — the handle of a value that lives in exactly one party.

Parity: fed/fed_object.py:18-80.  A handle knows its owner party, the sequence
id of the producing task, an output index and — in the owner party, or after a
receive — the local future (`Ref`) of the value.  Two pieces of per-object
bookkeeping implement "push at most once per destination" and "receive at most
once" (contract items 7 in SURVEY Appendix A).
"""
from __future__ import annotations

from typing import Set


class FedObjectSendingContext:
    """Which parties this object is being / has been pushed to."""

    __slots__ = ("_sent_to",)

    def __init__(self) -> None:
        self._sent_to: Set[str] = set()

    def mark_is_sending_to_party(self, target_party: str) -> None:
        self._sent_to.add(target_party)

    def was_sending_or_sent_to_party(self, target_party: str) -> bool:
        return target_party in self._sent_to


class FedObjectReceivingContext:
    __slots__ = ()


_NO_RECEIVING_CONTEXT = FedObjectReceivingContext()


class FedObject:
    __slots__ = ("_node_party", "_fed_task_id", "_ref", "_idx_in_task", "_sending_context", "_receiving_context")

    def __init__(self, node_party: str, fed_task_id: int, ref=None, idx_in_task: int = 0) -> None:
        self._node_party = node_party
        self._fed_task_id = fed_task_id
        self._ref = ref
        self._idx_in_task = idx_in_task
        self._sending_context = None  # allocated on the first push (most objects never leave)
        self._receiving_context = _NO_RECEIVING_CONTEXT

    # -- identity ------------------------------------------------------------
    def get_party(self) -> str:
        return self._node_party

    def get_fed_task_id(self) -> str:
        """``"<task seq id>#<output index>"`` — the upstream half of a message key."""
        return f"{self._fed_task_id}#{self._idx_in_task}"

    # -- local future ----------------------------------------------------------
    def get_ref(self):
        return self._ref

    # name kept for drop-in compatibility with code written against the reference
    get_ray_object_ref = get_ref

    def _cache_ref(self, ref) -> None:
        self._ref = ref

    _cache_ray_object_ref = _cache_ref

    # -- push bookkeeping --------------------------------------------------------
    def _mark_is_sending_to_party(self, target_party: str) -> None:
        ctx = self._sending_context
        if ctx is None:
            ctx = self._sending_context = FedObjectSendingContext()
        ctx.mark_is_sending_to_party(target_party)

    def _was_sending_or_sent_to_party(self, target_party: str) -> bool:
        ctx = self._sending_context
        return ctx is not None and ctx.was_sending_or_sent_to_party(target_party)

    def __repr__(self) -> str:
        where = "local" if self._ref is not None else "remote"
        return f"FedObject(party={self._node_party!r}, id={self.get_fed_task_id()}, {where})"
