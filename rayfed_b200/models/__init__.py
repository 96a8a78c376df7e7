"""Model families used by the federated examples and benchmarks.

* :mod:`.resnet` — ResNet-18/50 (channels-last, bf16) for FedAvg image training.
* :mod:`.llama` — Llama-3 style decoder with LoRA adapters (8B config + tiny
  configs for tests) for federated LoRA fine-tuning.
* :mod:`.mlp` — small MLP / split-learning halves for CPU tests.
"""
from .llama import LlamaConfig, LlamaLoRA, LoRALinear, llama_lora  # noqa: F401
from .mlp import MLP, SplitBottom, SplitTop  # noqa: F401
from .resnet import resnet18, resnet50  # noqa: F401
