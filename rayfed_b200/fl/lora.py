"""Federated LoRA on the B200 path.

Two aggregation rules, both one collective per round with no host hop:

* :class:`FedLoRAAdam` — the adapters of every party live in one flat bf16
  :class:`FlatBucket` in the symmetric heap; the fused aggregate kernel (K3)
  all-reduces them (weights *or* gradients), applies Adam/AdamW (FedAdam) on the
  sharded fp32 master and writes the new adapters to every party — optionally
  also a block-scaled fp8 copy.  This is BASELINE.json's "FedAvg Llama-3-8B LoRA
  fp8, fused all-reduce + Adam".
* :class:`FedLoRAExactMerge` — the *exact* FedAvg of the low-rank updates,
  ``W ← W + (alpha/r)·(1/P)·Σ_i B_i A_i``, computed by the tcgen05/TMEM kernel
  whose K dimension is spread over the parties (operands gathered over NVLink
  inside the kernel); adapters are then re-initialised (B = 0) for the next
  round.  Works on bf16 or block-scaled fp8 base weights.
"""
from __future__ import annotations

import torch

from ..models.llama import LlamaLoRA
from ..ops import kernels as K
from ..parallel.group import PartyGroup
from .bucket import FlatBucket
from .fedavg import FedAggregator


class FedLoRAAdam:
    def __init__(self, model: LlamaLoRA, group: PartyGroup, heap, channel=None, lr: float = 1e-3,
                 betas=(0.9, 0.999), weight_decay: float = 0.0, on_grads: bool = True, out_fp8: bool = False,
                 multicast: bool = True):
        self.model, self.group = model, group
        self.bucket = FlatBucket(model, heap if group.size > 1 else heap, name="lora", n_parties=group.size,
                                 multicast=multicast, grads_symmetric=on_grads and group.size > 1)
        kind = "adam" if on_grads else "fedadam"
        self.agg = FedAggregator(self.bucket, group, channel, kind=kind, lr=lr, betas=betas,
                                 weight_decay=weight_decay, out_fp8=out_fp8, heap=heap)

    def step(self, stream=None):
        """Collective: all-reduce + Adam + all-gather of the adapters (one kernel)."""
        self.agg.aggregate(stream)
        if self.agg.in_is_grad:
            self.bucket.flat_g.zero_()


class FedLoRAExactMerge:
    """Per adapted projection: symmetric [d_out, r] / [d_in, r] buffers + one K3 launch."""

    def __init__(self, model: LlamaLoRA, group: PartyGroup, heap, channel=None):
        self.model, self.group, self.heap, self.channel = model, group, heap, channel
        self.entries = []
        dev = next(model.parameters()).device
        for name, mod in model.lora_modules():
            assert mod.r % 16 == 0 and 16 <= mod.r * group.size <= 256, "K3 needs 16 <= parties*r <= 256"
            if group.size > 1:
                bb = heap.alloc(f"lora-B/{name}", mod.d_out * mod.r * 2)
                ab = heap.alloc(f"lora-At/{name}", mod.d_in * mod.r * 2)
                b_view = bb.tensor(torch.bfloat16, mod.d_out * mod.r).view(mod.d_out, mod.r)
                a_view = ab.tensor(torch.bfloat16, mod.d_in * mod.r).view(mod.d_in, mod.r)
                with torch.no_grad():
                    b_view.copy_(mod.lora_B)
                    a_view.copy_(mod.lora_At)
                mod.lora_B.data = b_view
                mod.lora_At.data = a_view
                b_ptrs, a_ptrs = list(bb.peer_ptrs), list(ab.peer_ptrs)
            else:
                b_ptrs, a_ptrs = [mod.lora_B.data_ptr()], [mod.lora_At.data_ptr()]
            self.entries.append((name, mod, b_ptrs, a_ptrs))
        self._dev = dev

    def _barrier(self, stream=None):
        if self.channel is not None and self.group.size > 1:
            self.channel.device_barrier(stream)

    @torch.no_grad()
    def merge(self, stream=None, reinit: bool = True):
        """Collective. Every party ends with the same merged base weights."""
        self._barrier(stream)  # all adapters final before anyone reads them
        for name, mod, b_ptrs, a_ptrs in self.entries:
            scale = mod.scaling / self.group.size
            if mod.base_dtype == "fp8":
                K.lora_merge(b_ptrs=b_ptrs, at_ptrs=a_ptrs, out_ptr=mod.weight_q.data_ptr(), m=mod.d_out,
                             n=mod.d_in, r=mod.r, scale=scale, w_base_ptr=mod.weight_q.data_ptr(),
                             w_base_scales_ptr=mod.weight_s.data_ptr(), out_scales_ptr=mod.weight_s.data_ptr(),
                             base_fp8=True, out_fp8=True, stream=stream)
            else:
                K.lora_merge(b_ptrs=b_ptrs, at_ptrs=a_ptrs, out_ptr=mod.weight.data_ptr(), m=mod.d_out,
                             n=mod.d_in, r=mod.r, scale=scale, w_base_ptr=mod.weight.data_ptr(), stream=stream)
        self._barrier(stream)  # everyone finished reading before adapters are reset
        if reinit:
            for _, mod, _, _ in self.entries:
                mod.lora_B.zero_()
