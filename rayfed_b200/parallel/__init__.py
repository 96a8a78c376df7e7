"""Cross-party parallelism building blocks.

* :mod:`.symmetric` — the symmetric heap: every party allocates the same buffer,
  exchanges handles, and ends up with a device pointer to every peer's copy
  (cudaIpc; CUDA-VMM + NVLS multicast when available).
* :mod:`.group` — ``PartyGroup``: ranks, exchange of small blobs between
  parties over the fed transport, torch.distributed or a filesystem rendezvous.
* Strategies: cross-party data parallelism (FedAvg / FedSGD: :mod:`rayfed_b200.fl`)
  and the split-learning pipeline cut (:mod:`.split`).
"""
from .group import FedExchange, FileExchange, PartyGroup, TorchDistExchange  # noqa: F401
from .symmetric import SymmetricBuffer, SymmetricHeap  # noqa: F401
