"""PartyGroup — who the parties are, and how they swap small blobs at set-up.

Three interchangeable exchanges (all implement ``all_gather(tag, blob)``):

* :class:`FedExchange` — over the fed transport itself with reserved keys; this
  is what ``fed.init`` uses, so no second rendezvous mechanism is needed;
* :class:`TorchDistExchange` — over an initialised ``torch.distributed`` group
  (bench.py under torchrun);
* :class:`FileExchange` — a directory on a shared filesystem (tests, tools).
"""
from __future__ import annotations

import os
import time
from typing import List, Sequence


class PartyGroup:
    def __init__(self, parties: Sequence[str], me: str, exchange, device_index: int = 0):
        self.parties = list(parties)
        self.me = me
        self.rank = self.parties.index(me)
        self.size = len(self.parties)
        self.exchange = exchange
        self.device_index = device_index
        self._tag_seq = 0

    def next_tag(self, prefix: str) -> str:
        self._tag_seq += 1
        return f"{prefix}:{self._tag_seq}"

    def all_gather(self, tag: str, blob: bytes) -> List[bytes]:
        """Every party contributes ``blob``; returns the blobs in party order."""
        if self.size == 1:
            return [blob]
        return self.exchange.all_gather(tag, blob, self)

    def barrier(self, tag: str = None):
        self.all_gather(tag or self.next_tag("barrier"), b"")


class FedExchange:
    """Blobs ride the already-running fed transport under reserved seq ids."""

    def all_gather(self, tag: str, blob: bytes, group: PartyGroup) -> List[bytes]:
        from ..proxy.barriers import recv, send

        for p in group.parties:
            if p != group.me:
                send(p, blob, f"__xchg__/{tag}/{group.me}", f"__xchg__/{tag}", check_sending=False)
        out = []
        for p in group.parties:
            if p == group.me:
                out.append(blob)
            else:
                out.append(recv(group.me, p, f"__xchg__/{tag}/{p}", f"__xchg__/{tag}").result(120.0))
        return out


class TorchDistExchange:
    def __init__(self, process_group=None):
        self._pg = process_group

    def all_gather(self, tag: str, blob: bytes, group: PartyGroup) -> List[bytes]:
        import torch.distributed as dist

        out = [None] * group.size
        dist.all_gather_object(out, (tag, blob), group=self._pg)
        assert all(t == tag for t, _ in out), f"exchange tag mismatch: {[t for t, _ in out]}"
        return [b for _, b in out]


class FileExchange:
    def __init__(self, directory: str, timeout_s: float = 120.0):
        self._dir = directory
        self._timeout = timeout_s
        os.makedirs(directory, exist_ok=True)

    def all_gather(self, tag: str, blob: bytes, group: PartyGroup) -> List[bytes]:
        safe = tag.replace("/", "_").replace(":", "_")
        mine = os.path.join(self._dir, f"{safe}.{group.me}")
        tmp = mine + ".tmp"
        with open(tmp, "wb") as f:
            f.write(blob)
        os.replace(tmp, mine)
        out = []
        deadline = time.time() + self._timeout
        for p in group.parties:
            path = os.path.join(self._dir, f"{safe}.{p}")
            while not os.path.exists(path):
                if time.time() > deadline:
                    raise TimeoutError(f"FileExchange: party {p} never published {tag}")
                time.sleep(0.002)
            with open(path, "rb") as f:
                out.append(f.read())
        return out
