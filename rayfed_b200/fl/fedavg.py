"""FedAvg on the B200 path: fused local optimizer + fused cross-party aggregate."""
from __future__ import annotations

import logging
import os
from typing import Callable, List, Optional, Sequence

import torch

from .. import _tracing
from ..ops import kernels as K
from ..parallel.group import FedExchange, PartyGroup
from ..parallel.symmetric import SymmetricHeap
from .bucket import FlatBucket

logger = logging.getLogger(__name__)

_KINDS = {"fedavg": K.AGG_AVG, "fedavgm": K.AGG_FEDAVGM, "fedsgd": K.AGG_SGD, "adam": K.AGG_ADAM,
          "fedadam": K.AGG_FEDADAM}


def device_context(parties: Optional[Sequence[str]] = None, me: Optional[str] = None, device: Optional[str] = None):
    """(group, heap, channel) of this party.

    Inside a fed job this is the DeviceChannel that ``fed.init`` created for the
    NVLink transport; stand-alone (tests, single party) a private one-party
    context is built.
    """
    from ..proxy.nvlink import nvlink_proxy

    ch = nvlink_proxy.current_channel()
    if ch is not None:
        return ch.group, ch.heap, ch
    dev = torch.device(device or f"cuda:{torch.cuda.current_device()}")
    group = PartyGroup(list(parties or ["solo"]), me or (parties[0] if parties else "solo"), FedExchange(),
                       device_index=dev.index or 0)
    if group.size != 1:
        raise RuntimeError("multi-party device_context needs the NVLink transport (fed.init with a device)")
    return group, SymmetricHeap(group), None


class LocalSGD:
    """Fused SGD(+momentum, +wd) step on a :class:`FlatBucket`; clears grads in the same pass."""

    def __init__(self, bucket: FlatBucket, lr: float, momentum: float = 0.9, weight_decay: float = 0.0):
        assert bucket.dtype == torch.bfloat16, "the fused local optimizer works on bf16 buckets"
        self.bucket = bucket
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.mom = torch.zeros(bucket.numel, dtype=torch.float32, device=bucket.flat_w.device)
        self._dev = bucket.flat_w.device.index or 0

    def step(self, grad_scale: float = 1.0, stream=None):
        b = self.bucket
        K.local_sgd(b.flat_w.data_ptr(), b.flat_g.data_ptr(), self.mom.data_ptr(), b.numel, lr=self.lr,
                    momentum=self.momentum, weight_decay=self.weight_decay, grad_scale=grad_scale, zero_grad=True,
                    stream=stream, device_index=self._dev)

    def reset_momentum(self):
        self.mom.zero_()


class FedAggregator:
    """One fused launch per round: all-reduce + scale + server optimizer + all-gather.

    ``kind``:
      ``fedavg``   w ← Σ wₖ·xₖ (xₖ = party k's weights)
      ``fedavgm``  server momentum on the pseudo-gradient (master − mean)
      ``fedsgd``   xₖ are gradients: SGD(+momentum, +wd) on the sharded fp32 master
      ``adam``     xₖ are gradients: AdamW on the sharded master
      ``fedadam``  AdamW on the pseudo-gradient
    Optimizer state is sharded: each party keeps fp32 master/moments only for the
    1/N of the elements it owns.
    """

    def __init__(self, bucket: FlatBucket, group: PartyGroup, channel=None, kind: str = "fedavg",
                 weights: Optional[Sequence[float]] = None, lr: float = 1.0, momentum: float = 0.0,
                 weight_decay: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-8, out_fp8: bool = False,
                 grads: bool = False, use_multicast: Optional[bool] = None, timeout_s: float = 120.0, heap=None):
        self.bucket, self.group, self.channel = bucket, group, channel
        self.kind = _KINDS[kind]
        self.kind_name = kind
        n = group.size
        self.weights = list(weights) if weights is not None else [1.0 / n] * n
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.betas, self.eps = betas, eps
        self.step_count = 0
        self.dev = bucket.flat_w.device
        self.shard_len = bucket.numel // n
        self.shard_start = group.rank * self.shard_len
        self.in_is_grad = grads or kind in ("fedsgd", "adam")
        self.src = bucket.flat_g if self.in_is_grad else bucket.flat_w
        self.master = self.mom = self.var = None
        if self.kind != K.AGG_AVG:
            sl = slice(self.shard_start, self.shard_start + self.shard_len)
            self.master = bucket.flat_w[sl].float().clone()
            self.mom = torch.zeros_like(self.master)
            if self.kind in (K.AGG_ADAM, K.AGG_FEDADAM):
                self.var = torch.zeros_like(self.master)
        self.out_fp8 = out_fp8
        self._timeout_ns = int(timeout_s * 1e9)  # device waits use %globaltimer (ns)
        self._host_gate = None  # decided at the first rendezvous from the driver's module-loading mode
        self._entry_ordered = False
        # A private signal pad per aggregator (symmetric when there are peers) and a
        # device-resident epoch: the launch is identical every round => graph-replayable.
        heap = heap or (channel.heap if channel is not None else None)
        self._ctl = torch.zeros(64, dtype=torch.int32, device=self.dev)  # [0]=epoch [1]=done
        self._status_host = torch.zeros(16, dtype=torch.int32).pin_memory()  # written by the kernel on timeout
        if n > 1:
            assert heap is not None, "multi-party aggregation needs the symmetric heap"
            self._pad = heap.alloc(f"aggpad/{bucket.name}/{kind}", 64 * 1024)
            sig_ptr, sig_peers = self._pad.local_ptr, list(self._pad.peer_ptrs)
            # arrival words of the pre-launch rendezvous: the last 64 bytes of the pad (per-CTA barrier
            # slots occupy 64 B per CTA from the start; grids never reach CTA 1023)
            self._arrive_off = 64 * 1024 - 64
        else:
            self._pad_t = torch.zeros(16384, dtype=torch.int32, device=self.dev)
            sig_ptr, sig_peers = self._pad_t.data_ptr(), [self._pad_t.data_ptr()]
        if self.in_is_grad:
            gb = getattr(bucket, "grad_buffer", None)
            if gb is None and n > 1:
                raise NotImplementedError(
                    "gradient aggregation across parties needs FlatBucket(grads_symmetric=True)")
            ins = gb.peer_ptrs if gb is not None else [bucket.flat_g.data_ptr()]
            mc = gb.mc_ptr if gb is not None else 0
        else:
            ins = bucket.weight_ptrs()
            mc = bucket.buffer.mc_ptr if bucket.buffer is not None else 0
        mc_out = bucket.buffer.mc_ptr if bucket.buffer is not None else 0
        if use_multicast is None:
            # measured on B200 (profiles/README.md): at 2 and 4 parties the P2P path is
            # faster than NVLS (N=4: 148 vs 161 us for 51 MB); NVLS halves the egress
            # bytes only asymptotically, so it is enabled from 8 parties.
            use_multicast = n >= 8
            env = os.environ.get("RAYFED_B200_AGG_MC")  # experiments: force NVLS on (1) / off (0)
            if env in ("0", "1"):
                use_multicast = env == "1"
        if use_multicast is False or n == 1:
            mc = mc_out = 0
        uniform = all(abs(w - self.weights[0]) < 1e-12 for w in self.weights)
        if not (mc and mc_out and uniform and not out_fp8):
            mc = mc_out = 0
        self._mc = mc
        self._scales = None
        out_ptrs, out_scales = bucket.weight_ptrs(), ()
        if out_fp8:
            fb = bucket.ensure_fp8_buffers(heap)
            out_ptrs, out_scales = fb[0], fb[1]
        base = self._ctl.data_ptr()
        self._plan = K.AggregatePlan(
            inputs=ins, outputs=out_ptrs, out_scales=out_scales, my_rank=group.rank, shard_start=self.shard_start,
            shard_len=self.shard_len, kind=self.kind, weights=self.weights, in_fp32=False, out_fp8=out_fp8,
            mc_in=mc, mc_out=mc_out,
            master_ptr=self.master.data_ptr() if self.master is not None else 0,
            mom_ptr=self.mom.data_ptr() if self.mom is not None else 0,
            var_ptr=self.var.data_ptr() if self.var is not None else 0,
            lr=lr, momentum=momentum, weight_decay=weight_decay, beta1=betas[0], beta2=betas[1], eps=eps,
            sig_ptr=sig_ptr, sig_peers=sig_peers, epoch_dev_ptr=base, agg_done_ptr=base + 4, status_ptr=self._status_host.data_ptr(),
            timeout_ns=self._timeout_ns, device_index=self.dev.index or 0,
        )

    def _rendezvous(self, stream):
        """Everybody reaches the collective before anybody launches the spinning kernel (stream
        memory operations: no SM waits).  Skipped under CUDA-graph capture, where the in-kernel
        entry barrier alone orders the parties (the caller then owns the no-starvation argument)."""
        ordered = self.group.size > 1 and not torch.cuda.is_current_stream_capturing()
        if ordered != self._entry_ordered:
            self._plan.set_entry_ordered(ordered)
            self._entry_ordered = ordered
        if ordered:
            off = self._arrive_off
            if self._host_gate is None:
                mode = os.environ.get("RAYFED_B200_DEVICE_WAIT", "auto")  # auto | stream | host
                try:
                    lazy = K.cuda_ext().module_loading_mode() != "eager"
                except Exception:  # noqa: BLE001
                    lazy = True
                self._host_gate = mode == "host" or (mode == "auto" and lazy)
            K.rendezvous(arrive_peers=[p + off for p in self._pad.peer_ptrs], arrive_local=self._pad.local_ptr + off,
                         my_rank=self.group.rank, value=self.step_count, stream=stream, host_gate=self._host_gate,
                         timeout_s=max(600.0, self._timeout_ns / 1e9))

    def aggregate(self, stream=None, lr: Optional[float] = None):
        """Enqueue the fused aggregate on ``stream`` (default: current). Collective."""
        self.step_count += 1
        self._rendezvous(stream)
        tl = _tracing._TRACER
        if tl is None or torch.cuda.is_current_stream_capturing():
            self._plan.launch(0, self.step_count, lr, stream)
            return
        s = stream if stream is not None else torch.cuda.current_stream()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(s)
        self._plan.launch(0, self.step_count, lr, stream)
        ev1.record(s)
        tl.device_span("fed_aggregate", ev0, ev1, round=self.step_count, shard_elems=self.shard_len)

    def phase_breakdown(self, iters: int = 5):
        """Per-phase device time of the pipelined FedAvg kernel (collective; plain ``fedavg`` of bf16
        weights only): median over ``iters`` launches of the max over CTAs of {entry barrier, reduce +
        gather loop, exit barrier} and the span first-start .. last-exit, in microseconds."""
        ts = torch.zeros(4 * 1024, dtype=torch.int64, device=self.dev)
        self._plan.set_phase_ts(ts.data_ptr())
        rows = []
        try:
            for _ in range(iters):
                ts.zero_()
                self.aggregate()
                torch.cuda.current_stream().synchronize()
                t = ts.view(-1, 4)
                t = t[t[:, 3] > 0].double()
                if t.numel() == 0:
                    return None
                rows.append({"ctas": int(t.shape[0]),
                             "entry_barrier_us": float((t[:, 1] - t[:, 0]).max()) / 1e3,
                             "loop_us": float((t[:, 2] - t[:, 1]).max()) / 1e3,
                             "loop_us_median_cta": float((t[:, 2] - t[:, 1]).median()) / 1e3,
                             "exit_barrier_us": float((t[:, 3] - t[:, 2]).max()) / 1e3,
                             "span_us": float(t[:, 3].max() - t[:, 0].min()) / 1e3})
        finally:
            self._plan.set_phase_ts(0)
        rows.sort(key=lambda r: r["span_us"])
        return rows[len(rows) // 2]

    def status(self) -> int:
        """0 = healthy; 2 = an in-kernel party barrier timed out."""
        return int(self._status_host[0])

    @property
    def uses_multicast(self) -> bool:
        return bool(self._mc)


def aggregate_reference(worker_weights: List, fed, server_party: str, parties: Sequence[str]):
    """Reference-style FedAvg round through the public DAG API (the baseline):
    every party's weights are pushed to ``server_party``, averaged there, and the
    mean is pushed back to every party."""

    @fed.remote
    def _mean(*ws):
        return torch.stack([w.float() for w in ws]).mean(0).to(ws[0].dtype)

    return _mean.party(server_party).remote(*worker_weights)


class FedAvgWorker:
    """Actor body: one party's model, optimizer, data and aggregator.

    Typical use (same script in every party)::

        Worker = fed.remote(FedAvgWorker)
        workers = {p: Worker.party(p).remote(model_fn, parties=parties, party=p, ...) for p in parties}
        for r in range(rounds):
            losses = [workers[p].local_round.remote(steps) for p in parties]
            for p in parties: workers[p].aggregate.remote()
            fed.get(losses)
    """

    def __init__(self, model_fn: Callable[[], torch.nn.Module], parties: Sequence[str], party: str,
                 batch_fn: Optional[Callable] = None, lr: float = 0.05, momentum: float = 0.9,
                 weight_decay: float = 5e-5, kind: str = "fedavg", server_lr: float = 1.0,
                 server_momentum: float = 0.0, channels_last: bool = True, multicast: bool = True,
                 seed: int = 0, loss_fn=None, use_cuda_graph: bool = False, fused_aggregate: bool = True):
        if fused_aggregate:
            group, heap, channel = device_context(parties, party)
        else:
            # weights are exchanged through the DAG API instead (reference-style aggregation):
            # plain local bucket, no symmetric heap, no aggregate kernel
            dev_idx = torch.cuda.current_device()
            group = PartyGroup([party], party, FedExchange(), device_index=dev_idx)
            heap, channel = None, None
        self.group, self.heap, self.channel = group, heap, channel
        dev = torch.device(f"cuda:{group.device_index}")
        torch.cuda.set_device(dev)
        torch.manual_seed(seed)  # identical initial weights in every party
        model = model_fn().to(dev)
        if channels_last:
            model = model.to(memory_format=torch.channels_last)
        self.model = model.to(torch.bfloat16)
        self.bucket = FlatBucket(self.model, heap, name="fedavg", n_parties=group.size, multicast=multicast)
        self.opt = LocalSGD(self.bucket, lr, momentum, weight_decay)
        self.agg = None
        if fused_aggregate:
            self.agg = FedAggregator(self.bucket, group, channel, kind=kind, lr=server_lr,
                                     momentum=server_momentum, heap=heap)
        self.loss_fn = loss_fn or torch.nn.functional.cross_entropy
        self.batch_fn = batch_fn
        self.channels_last = channels_last
        self.dev = dev
        self.rounds = 0
        self._graph = None
        self._use_graph = use_cuda_graph
        self._static = None

    # -- one local optimisation step ------------------------------------------------
    def _train_step(self, x, y):
        out = self.model(x)
        loss = self.loss_fn(out.float(), y)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def local_step(self, x, y):
        """x, y: host (pinned) or device tensors."""
        x = x.to(self.dev, non_blocking=True)
        y = y.to(self.dev, non_blocking=True)
        if x.dtype != torch.bfloat16:
            x = x.to(torch.bfloat16)
        if self.channels_last and x.dim() == 4:
            x = x.contiguous(memory_format=torch.channels_last)
        if not self._use_graph:
            return self._train_step(x, y)
        return self._graphed_step(x, y)

    def _graphed_step(self, x, y):
        if self._graph is None:
            self._static = (x.clone(), y.clone())
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):  # warm-up outside capture (cudnn autotune, allocator)
                    self._train_step(*self._static)
            torch.cuda.current_stream().wait_stream(s)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):  # other runtime threads keep issuing CUDA calls
                self._static_loss = self._train_step(*self._static)
        self._static[0].copy_(x, non_blocking=True)
        self._static[1].copy_(y, non_blocking=True)
        self._graph.replay()
        K.launch_counter.add("fed_local_sgd", 1)  # the captured step launches it once per replay
        return self._static_loss

    def local_round(self, steps: int = 1, batches=None):
        """``steps`` local SGD steps; returns the last loss as a 0-d device tensor."""
        loss = None
        for i in range(steps):
            x, y = batches[i] if batches is not None else self.batch_fn(self.rounds, i)
            loss = self.local_step(x, y)
        self.rounds += 1
        if self._use_graph and loss is not None:
            # the graph's output tensor is overwritten by the next replay; hand out a snapshot
            # (ordered on this stream) so a reader one round later still sees THIS round's loss
            loss = loss.clone()
        return loss

    def aggregate(self):
        self.agg.aggregate()
        return True

    def loss_value(self, loss):
        return float(loss)

    def weights_checksum(self):
        return float(self.bucket.flat_w.float().sum())

    def get_flat_weights(self):
        return self.bucket.flat_w.clone()

    def set_flat_weights(self, w):
        self.bucket.flat_w.copy_(w)
        return True
