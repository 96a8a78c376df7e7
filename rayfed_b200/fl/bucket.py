"""FlatBucket — model parameters as views of one flat buffer in the symmetric heap."""
from __future__ import annotations

from typing import List, Optional

import torch

from ..parallel.symmetric import SymmetricBuffer, SymmetricHeap


class FlatBucket:
    """Re-homes ``model``'s parameters into a flat ``dtype`` buffer.

    * the weight buffer is a :class:`SymmetricBuffer` (peers address it directly;
      the aggregate kernel updates it in place),
    * the gradient buffer is a local flat tensor; each ``param.grad`` is a view of
      it, so backward accumulates straight into the bucket,
    * the length is padded to a multiple of ``32 * n_parties`` elements so every
      party owns an equal, 32-aligned shard (fp8 block scales never straddle
      shards).

    4-D parameters keep ``channels_last`` strides when the model uses them.
    """

    def __init__(self, model: torch.nn.Module, heap: Optional[SymmetricHeap], name: str = "model",
                 dtype=torch.bfloat16, n_parties: int = 1, multicast: bool = False, device=None,
                 grads_symmetric: bool = False):
        params = [p for p in model.parameters() if p.requires_grad]
        self.name = name
        self.params: List[torch.nn.Parameter] = params
        self.dtype = dtype
        self.n_parties = n_parties
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 7) // 8 * 8  # keep every tensor 16-byte aligned in bf16
        align = 32 * max(1, n_parties)
        self.numel = (total + align - 1) // align * align
        self.offsets = offs
        itemsize = torch.empty((), dtype=dtype).element_size()
        if heap is not None:
            self.buffer: Optional[SymmetricBuffer] = heap.alloc(f"bucket/{name}", self.numel * itemsize,
                                                                 multicast=multicast)
            self.flat_w = self.buffer.tensor(dtype, self.numel)
            device = self.flat_w.device
        else:
            self.buffer = None
            device = device or params[0].device
            self.flat_w = torch.zeros(self.numel, dtype=dtype, device=device)
        self.grad_buffer = None
        if grads_symmetric and heap is not None:
            # gradients readable by peers (FedSGD / cross-party Adam on averaged grads)
            self.grad_buffer = heap.alloc(f"bucket-grad/{name}", self.numel * itemsize, multicast=multicast)
            self.flat_g = self.grad_buffer.tensor(dtype, self.numel)
        else:
            self.flat_g = torch.zeros(self.numel, dtype=dtype, device=device)
        self._fp8 = None
        with torch.no_grad():
            for p, off in zip(params, offs):
                n = p.numel()
                cl = p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous()
                w_view = self._view(self.flat_w, off, p, cl)
                w_view.copy_(p.detach().to(device=device, dtype=dtype))
                p.data = w_view
                p.grad = self._view(self.flat_g, off, p, cl)
        self.shard_len = self.numel // max(1, n_parties)

    @staticmethod
    def _view(flat, off, p, channels_last):
        n = p.numel()
        if channels_last:
            o, i, h, w = p.shape
            return flat[off: off + n].view(o, h, w, i).permute(0, 3, 1, 2)
        return flat[off: off + n].view(p.shape)

    @property
    def nbytes(self) -> int:
        return self.numel * self.flat_w.element_size()

    def weight_ptrs(self) -> List[int]:
        return self.buffer.peer_ptrs if self.buffer is not None else [self.flat_w.data_ptr()]

    def ensure_fp8_buffers(self, heap):
        """Block-scaled e4m3 copy of the weights (+ UE8M0 scale per 32) written by the
        fp8 aggregate; symmetric so every party receives the re-quantised weights."""
        if self._fp8 is None:
            if heap is not None and self.n_parties > 1:
                pay = heap.alloc(f"bucket-fp8/{self.name}", self.numel)
                sc = heap.alloc(f"bucket-fp8s/{self.name}", self.numel // 32)
                self.fp8_payload = pay.tensor(torch.uint8, self.numel)
                self.fp8_scales = sc.tensor(torch.uint8, self.numel // 32)
                self._fp8 = (list(pay.peer_ptrs), list(sc.peer_ptrs))
            else:
                self.fp8_payload = torch.zeros(self.numel, dtype=torch.uint8, device=self.flat_w.device)
                self.fp8_scales = torch.zeros(self.numel // 32, dtype=torch.uint8, device=self.flat_w.device)
                self._fp8 = ([self.fp8_payload.data_ptr()], [self.fp8_scales.data_ptr()])
        return self._fp8

    def dequantized_fp8(self):
        """fp32 view of the fp8 weights (test / debugging helper)."""
        q = self.fp8_payload.view(torch.float8_e4m3fn).float().view(-1, 32)
        s = torch.exp2(self.fp8_scales.float() - 127.0).view(-1, 1)
        return (q * s).view(-1)

    def zero_grad(self):
        self.flat_g.zero_()
