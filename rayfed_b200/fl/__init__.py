"""Federated-learning layer: cross-party data parallelism as a first-class op.

The reference has no FL primitive — FedAvg is user code that gathers weights to
one party, averages and redistributes them through the gRPC proxy
(fed/tests/test_fed_get.py:45-47,71-78).  Here:

* :class:`FlatBucket` keeps every parameter of a model as a view into one flat
  bf16 buffer that lives in the **symmetric heap**, so peers can read/write it;
* :class:`FedAggregator` launches the fused aggregate kernel (K2/K3): all-reduce
  + 1/N + server optimizer + all-gather in one launch per party, no host hop;
* :class:`LocalSGD` is the fused local optimizer on the flat bucket;
* :class:`FedAvgWorker` packages a model + data + the above as an actor body
  for ``@fed.remote``;
* :func:`aggregate_reference` is the reference-style aggregation (push weights
  to one party, mean, broadcast) used as the measured baseline.
"""
from .bucket import FlatBucket  # noqa: F401
from .lora import FedLoRAAdam, FedLoRAExactMerge  # noqa: F401
from .fedavg import FedAggregator, FedAvgWorker, LocalSGD, aggregate_reference, device_context  # noqa: F401
