"""Split learning — the pipeline-parallel-like cut across parties.

The reference names split learning only as an (empty) tutorial
(docs/source/tutorials/split_learning_demo.ipynb); the pattern is expressible
with the push primitive alone: the *bottom* party holds the features and the
lower layers, the *top* party holds the labels and the head; activations flow
forward and their gradients flow back as FedObjects (over NVLink when the
parties have GPUs).  These two actor bodies implement exactly that.
"""
from __future__ import annotations

import torch


class SplitBottomWorker:
    def __init__(self, model_fn, lr=0.05, device=None, seed=0):
        torch.manual_seed(seed)
        self.dev = torch.device(device) if device else torch.device("cpu")
        self.model = model_fn().to(self.dev)
        self.opt = torch.optim.SGD(self.model.parameters(), lr=lr)
        self._h = None

    def forward(self, x):
        self.opt.zero_grad(set_to_none=True)
        self._h = self.model(x.to(self.dev))
        return self._h.detach()

    def backward(self, grad_h):
        self._h.backward(grad_h.to(self.dev))
        self.opt.step()
        self._h = None
        return True


class SplitTopWorker:
    def __init__(self, model_fn, lr=0.05, device=None, seed=1, loss_fn=None):
        torch.manual_seed(seed)
        self.dev = torch.device(device) if device else torch.device("cpu")
        self.model = model_fn().to(self.dev)
        self.opt = torch.optim.SGD(self.model.parameters(), lr=lr)
        self.loss_fn = loss_fn or torch.nn.functional.cross_entropy

    def step(self, h, y):
        """Forward on the received activations, backward, return (loss, dL/dh)."""
        h = h.to(self.dev).detach().requires_grad_(True)
        self.opt.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model(h), y.to(self.dev))
        loss.backward()
        self.opt.step()
        return float(loss.detach()), h.grad.detach()
