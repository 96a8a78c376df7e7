"""Thin, allocation-free launch wrappers over ``_fedcuda``.

Pointers are passed as integers; the caller owns lifetime and stream ordering.
``launch_counter`` counts every kernel launched through this module (bench.py
reports it as ``gpu_launches``).
"""
from __future__ import annotations

import importlib
import threading
from typing import Optional, Sequence

from .. import _build

MODE_COPY = 0
MODE_F32_TO_BF16 = 1
MODE_BF16_TO_FP8 = 2
MODE_BF16_SCALE = 3
MODE_FP8_TO_BF16 = 4
MODE_BF16_TO_F32 = 5

AGG_AVG = 0
AGG_FEDAVGM = 1
AGG_SGD = 2
AGG_ADAM = 3
AGG_FEDADAM = 4

_ext = None
_ext_lock = threading.Lock()
_sm_count = {}


class _LaunchCounter:
    def __init__(self):
        self.count = 0
        self.by_kernel = {}

    def add(self, name: str, n: int = 1):
        self.count += n
        self.by_kernel[name] = self.by_kernel.get(name, 0) + n

    def reset(self):
        self.count = 0
        self.by_kernel = {}


launch_counter = _LaunchCounter()


def cuda_ext():
    """Load (building in-tree if needed) the CUDA extension; raise loudly if absent."""
    global _ext
    if _ext is None:
        with _ext_lock:
            if _ext is None:
                _build.ensure_cuda()
                _ext = importlib.import_module("rayfed_b200._fedcuda")
    return _ext


def sm_count(device_index: int = 0) -> int:
    if device_index not in _sm_count:
        _sm_count[device_index] = cuda_ext().sm_count(device_index)
    return _sm_count[device_index]


def _stream_handle(stream) -> int:
    if stream is None:
        import torch

        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream


def _blocks_for(nbytes: int, device_index: int, bytes_per_cta: int = 64 * 1024, max_per_sm: int = 4) -> int:
    sms = sm_count(device_index)
    want = max(1, (nbytes + bytes_per_cta - 1) // bytes_per_cta)
    return int(min(want, sms * max_per_sm))


def push(src_ptr: int, dst_ptr: int, n: int, *, mode: int = MODE_COPY, scale: float = 1.0, dst_scales_ptr: int = 0,
         flag_ptr: int = 0, flag_value: int = 0, done_counter_ptr: int, stream=None, device_index: int = 0,
         use_bulk: Optional[bool] = None, blocks: Optional[int] = None, desc_dst: int = 0, desc: bytes = b"") -> None:
    """K1. ``n`` is bytes for MODE_COPY, elements otherwise.  ``desc`` (64 bytes) is stored to
    ``desc_dst`` in peer memory by the last CTA, before the release flag."""
    nbytes = n if mode == MODE_COPY else n * (4 if mode == MODE_F32_TO_BF16 else 2)
    if use_bulk is None:
        use_bulk = mode == MODE_COPY and nbytes >= (1 << 20) and nbytes % 16 == 0 and _bulk_default()
    if blocks is None:
        blocks = sm_count(device_index) if use_bulk else _blocks_for(nbytes, device_index)
        if use_bulk:
            blocks = int(min(blocks, max(1, (nbytes + 32 * 1024 - 1) // (32 * 1024))))
    cuda_ext().push(src_ptr, dst_ptr, dst_scales_ptr, n, float(scale), flag_ptr, flag_value, done_counter_ptr, mode,
                    blocks, bool(use_bulk), _stream_handle(stream), desc_dst, desc)
    launch_counter.add("fed_push_bulk" if use_bulk else "fed_push")


def push_multi(segs: Sequence, dst_ptr: int, *, flag_ptr: int = 0, flag_value: int = 0, done_counter_ptr: int,
               stream=None, device_index: int = 0, blocks: Optional[int] = None, desc_dst: int = 0,
               desc: bytes = b"") -> None:
    """K1 for a whole object: ``segs`` = [(src_ptr, dst_offset, nbytes)] (<= 48, 16-byte aligned)
    land in one block at ``dst_ptr`` with one launch, one descriptor, one flag."""
    total = sum(s[2] for s in segs)
    if blocks is None:
        blocks = _blocks_for(total, device_index)
    cuda_ext().push_multi([tuple(int(x) for x in s) for s in segs], dst_ptr, flag_ptr, flag_value, done_counter_ptr,
                          blocks, _stream_handle(stream), desc_dst, desc)
    launch_counter.add("fed_push_multi")


def push_mc(src_ptr: int, mc_dst_ptr: int, nbytes: int, *, flag_ptrs: Sequence[int], flag_value: int,
            done_counter_ptr: int, stream=None, device_index: int = 0, blocks: Optional[int] = None,
            desc_dst: int = 0, desc: bytes = b"") -> None:
    """K1b: multicast push — one ``multimem.st`` stream lands in every party's copy of the
    destination; then the descriptor (multicast too) and the release flag of each listed peer."""
    if blocks is None:
        blocks = _blocks_for(nbytes, device_index)
    cuda_ext().push_mc(src_ptr, mc_dst_ptr, nbytes, list(flag_ptrs), flag_value, done_counter_ptr, blocks,
                       _stream_handle(stream), desc_dst, desc)
    launch_counter.add("fed_push_mc")


_BULK_DEFAULT = None


def _bulk_default() -> bool:
    global _BULK_DEFAULT
    if _BULK_DEFAULT is None:
        import os

        _BULK_DEFAULT = os.environ.get("RAYFED_B200_PUSH_BULK", "0") == "1"
    return _BULK_DEFAULT


def unpack(src_ptr: int, dst_ptr: int, n: int, *, mode: int = MODE_COPY, scale: float = 1.0, scales_ptr: int = 0,
           flag_ptr: int, flag_value: int, done_counter_ptr: int, credit_ptr: int = 0, credit_value: int = 0,
           status_ptr: int = 0, timeout_ns: int = 0, stream=None, device_index: int = 0,
           blocks: Optional[int] = None) -> None:
    """K4 + unpack: wait for the flag on the device, undo the wire format, return credit."""
    nbytes = n if mode == MODE_COPY else n * 2
    if blocks is None:
        blocks = _blocks_for(nbytes, device_index)
    cuda_ext().unpack(src_ptr, dst_ptr, scales_ptr, n, float(scale), flag_ptr, flag_value, done_counter_ptr,
                      credit_ptr, credit_value, status_ptr, timeout_ns, mode, blocks, _stream_handle(stream))
    launch_counter.add("fed_unpack")


def aggregate(*, inputs: Sequence[int], outputs: Sequence[int], my_rank: int, shard_start: int, shard_len: int,
              kind: int = AGG_AVG, weights: Optional[Sequence[float]] = None, in_fp32: bool = False,
              out_fp8: bool = False, out_scales: Sequence[int] = (), mc_in: int = 0, mc_out: int = 0,
              master_ptr: int = 0, mom_ptr: int = 0, var_ptr: int = 0, lr: float = 1.0, momentum: float = 0.0,
              weight_decay: float = 0.0, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, step: int = 1,
              sig_ptr: int, sig_peers: Sequence[int], epoch: int, status_ptr: int = 0, timeout_ns: int = 0,
              stream=None, device_index: int = 0, blocks: Optional[int] = None) -> None:
    """K2/K3. Every party launches this with the same arguments (SPMD)."""
    n = len(inputs)
    if weights is None:
        weights = [1.0 / n] * n
    if blocks is None:
        blocks = int(min(sm_count(device_index), max(1, (shard_len // 8 + 511) // 512)))
    bias_c1 = 1.0 / (1.0 - beta1 ** step) if beta1 < 1.0 else 1.0
    bias_c2 = 1.0 / (1.0 - beta2 ** step) if beta2 < 1.0 else 1.0
    cuda_ext().aggregate(list(inputs), list(outputs), list(out_scales), mc_in, mc_out, master_ptr, mom_ptr, var_ptr,
                         shard_start, shard_len, my_rank, kind, in_fp32, out_fp8, [float(w) for w in weights],
                         float(lr), float(momentum), float(weight_decay), float(beta1), float(beta2), float(eps),
                         float(bias_c1), float(bias_c2), sig_ptr, list(sig_peers), epoch, status_ptr, timeout_ns,
                         blocks, _stream_handle(stream))
    launch_counter.add("fed_aggregate")


class AggregatePlan:
    """Pre-baked K2/K3 launch: argument marshalling happens once, a round costs one
    pybind call.  With ``epoch_dev_ptr`` the epoch lives on the device and the launch
    is CUDA-graph replayable."""

    def __init__(self, *, inputs, outputs, my_rank, shard_start, shard_len, kind=AGG_AVG, weights=None,
                 in_fp32=False, out_fp8=False, out_scales=(), mc_in=0, mc_out=0, master_ptr=0, mom_ptr=0, var_ptr=0,
                 lr=1.0, momentum=0.0, weight_decay=0.0, beta1=0.9, beta2=0.999, eps=1e-8, sig_ptr, sig_peers,
                 epoch_dev_ptr=0, agg_done_ptr=0, status_ptr=0, timeout_ns=0, device_index=0, blocks=None):
        n = len(inputs)
        if weights is None:
            weights = [1.0 / n] * n
        if blocks is None:
            blocks = int(min(sm_count(device_index), max(1, (shard_len // 8 + 511) // 512)))
        self.lr, self.beta1, self.beta2 = float(lr), float(beta1), float(beta2)
        self._plan = cuda_ext().AggPlan(list(inputs), list(outputs), list(out_scales), mc_in, mc_out, master_ptr,
                                        mom_ptr, var_ptr, shard_start, shard_len, my_rank, kind, bool(in_fp32),
                                        bool(out_fp8), [float(w) for w in weights], float(lr), float(momentum),
                                        float(weight_decay), float(beta1), float(beta2), float(eps), sig_ptr,
                                        list(sig_peers), epoch_dev_ptr, agg_done_ptr, status_ptr, timeout_ns,
                                        blocks)

    def set_entry_ordered(self, on: bool):
        """The caller ran the pre-launch rendezvous: the kernel may skip its entry barrier."""
        self._plan.set_entry_ordered(bool(on))

    def set_force_generic(self, on: bool):
        """Experiments: route plain FedAvg through the generic aggregate kernel."""
        self._plan.set_force_generic(bool(on))

    def set_grid_cap(self, n: int):
        """Experiments: cap the pipelined FedAvg kernel's grid (0 = default, 2 CTAs per SM)."""
        self._plan.set_grid_cap(int(n))

    def set_phase_ts(self, ptr: int):
        """Debug: per-CTA %globaltimer stamps of the pipelined FedAvg kernel (see aggregate.cu)."""
        self._plan.set_phase_ts(int(ptr))

    def launch(self, epoch: int = 0, step: int = 1, lr: Optional[float] = None, stream=None):
        b1, b2 = self.beta1, self.beta2
        c1 = 1.0 / (1.0 - b1 ** step) if b1 < 1.0 else 1.0
        c2 = 1.0 / (1.0 - b2 ** step) if b2 < 1.0 else 1.0
        self._plan.launch(epoch, self.lr if lr is None else float(lr), c1, c2, _stream_handle(stream))
        launch_counter.add("fed_aggregate")


_gate_streams = {}


def rendezvous(*, arrive_peers: Sequence[int], arrive_local: int, my_rank: int, value: int, stream=None,
               host_gate: bool = False, timeout_s: float = 600.0) -> None:
    """SM-free N-party rendezvous in front of a collective kernel: one tiny kernel tells every peer
    "party `my_rank` is here" (``st.release.sys`` into word ``my_rank`` of each peer's arrival array),
    then the stream parks on its own arrival words with ``cuStreamWaitValue32`` — no SM is occupied
    while waiting, so the collective kernel (which spins in its in-kernel barriers) is not launched
    before every party has reached it and cannot starve local kernels a late party depends on."""
    ext = cuda_ext()
    sh = _stream_handle(stream)
    ext.signal_peers([int(p) for p in arrive_peers], int(my_rank), int(value) & 0xFFFFFFFF, sh)
    launch_counter.add("fed_signal")
    if host_gate:
        # lazy CUDA module loading: a parked stream wait would block first-time kernel launches of
        # this process (see proxy/nvlink/channel.py), so the host waits instead — tiny D2H reads of
        # the arrival words on a private stream; the stream waits below are then already satisfied
        import time as _time

        import torch

        dev = torch.cuda.current_device()
        gs = _gate_streams.get(dev)
        if gs is None:
            gs = _gate_streams[dev] = ext.stream_create(0)
        want = int(value) & 0xFFFFFFFF
        deadline = _time.monotonic() + timeout_s
        pause = 5e-6
        n = len(arrive_peers)
        while True:
            words = ext.read_u64s(int(arrive_local), (n + 1) // 2, gs)
            vals = [(words[r // 2] >> (32 * (r % 2))) & 0xFFFFFFFF for r in range(n)]
            if all(r == my_rank or ((vals[r] - want) & 0xFFFFFFFF) < 0x80000000 for r in range(n)):
                break
            if _time.monotonic() > deadline:
                raise TimeoutError("a party never reached the collective (host-gated rendezvous)")
            _time.sleep(pause)
            pause = min(pause * 1.5, 2e-4)
    for r in range(len(arrive_peers)):
        if r != my_rank:
            ext.stream_wait_value32(sh, int(arrive_local) + 4 * r, int(value) & 0xFFFFFFFF)


def barrier(*, n_parties: int, my_rank: int, sig_ptr: int, sig_peers: Sequence[int], epoch: int, status_ptr: int = 0,
            timeout_ns: int = 0, stream=None) -> None:
    """K5."""
    cuda_ext().barrier(n_parties, my_rank, sig_ptr, list(sig_peers), epoch, status_ptr, timeout_ns,
                       _stream_handle(stream))
    launch_counter.add("fed_barrier")


_lora_ws = {}


def _lora_workspace(m: int, n: int, kpad: int, stream_handle: int):
    """K-major concatenated operands [m, kpad] / [n, kpad] for the staged LoRA merge,
    cached per (device, shape, stream) so graph replays and repeated rounds reuse them."""
    import torch

    key = (torch.cuda.current_device(), m, n, kpad, stream_handle)
    ws = _lora_ws.get(key)
    if ws is None:
        ws = (torch.zeros(m, kpad, dtype=torch.bfloat16, device="cuda"),
              torch.zeros(n, kpad, dtype=torch.bfloat16, device="cuda"))
        _lora_ws[key] = ws
    return ws


def lora_merge(*, b_ptrs: Sequence[int], at_ptrs: Sequence[int], out_ptr: int, m: int, n: int, r: int,
               scale: Optional[float] = None, weights: Optional[Sequence[float]] = None, w_base_ptr: int = 0,
               w_base_scales_ptr: int = 0, out_scales_ptr: int = 0, base_fp8: bool = False, out_fp8: bool = False,
               stream=None, use_tma: bool = True, impl: Optional[str] = None) -> None:
    """K3 GEMM (tcgen05/TMEM): ``out = W_base + scale * sum_i w_i * B_i @ A_i`` with the
    operands gathered from every party's adapter buffers. ``B_i``: [m, r]; ``A_i^T``: [n, r].

    ``impl``: ``"staged3"`` (gather once + persistent TMA/tcgen05 pipeline with separate 3-deep W /
    2-deep operand rings; the default where it applies: bf16 base/out, parties*r <= 128),
    ``"staged"`` (the same with one shared 2-stage ring), ``"tma"`` (per-tile gather, TMA epilogue),
    ``"direct"`` (per-tile gather, register epilogue; the only one for fp8 base/out).
    ``RAYFED_B200_LORA_IMPL`` overrides the default."""
    import os

    p = len(b_ptrs)
    if scale is None:
        scale = 1.0 / p
    if r % 16 or not (16 <= p * r <= 256):
        raise ValueError(f"lora_merge needs r % 16 == 0 and 16 <= parties*r <= 256 (got r={r}, parties={p})")
    if n % 32:
        raise ValueError("lora_merge needs n % 32 == 0")
    if impl is None:
        # staged3 (decoupled 3-deep W ring) validated on hardware in round 2: bit-identical to
        # staged, 21.4 vs 25.1 us at 4096^2 and 49.5 vs 63.2 us at 4096x14336 (cuBLAS addmm: 18.2 / 56.0)
        impl = os.environ.get("RAYFED_B200_LORA_IMPL", "staged3") if use_tma else "direct"
    if impl not in ("staged", "staged3", "tma", "direct"):
        raise ValueError(f"unknown lora_merge impl {impl!r}")
    deep = impl == "staged3"  # decoupled W / operand rings (opt-in until measured on hardware)
    if deep:
        impl = "staged"
    if impl == "staged" and (base_fp8 or out_fp8 or p * r > 128):
        impl, deep = "tma", False
    sh = _stream_handle(stream)
    bcat = atcat = kpad = 0
    if impl == "staged":
        kpad = max((p * r + 63) // 64 * 64, int(os.environ.get("RAYFED_B200_LORA_KPAD_MIN", "0")))
        ws = _lora_workspace(m, n, kpad, sh)
        bcat, atcat = ws[0].data_ptr(), ws[1].data_ptr()
    cuda_ext().lora_merge(list(b_ptrs), list(at_ptrs), w_base_ptr, w_base_scales_ptr, out_ptr, out_scales_ptr, m, n,
                          r, float(scale), [float(w) for w in (weights or [1.0] * p)], bool(base_fp8),
                          bool(out_fp8), sh, impl != "direct", bcat, atcat, kpad, deep)
    if impl == "staged":
        launch_counter.add("fed_lora_gather")
        launch_counter.add("fed_lora_gemm")
    else:
        launch_counter.add("fed_lora_merge")


def local_sgd(w_ptr: int, g_ptr: int, mom_ptr: int, n: int, *, lr: float, momentum: float = 0.0,
              weight_decay: float = 0.0, grad_scale: float = 1.0, zero_grad: bool = True, stream=None,
              device_index: int = 0, blocks: Optional[int] = None) -> None:
    if n % 8:
        raise ValueError("local_sgd needs a flat bucket whose length is a multiple of 8")
    if blocks is None:
        blocks = int(min(sm_count(device_index) * 8, max(1, (n // 8 + 511) // 512)))
    cuda_ext().local_sgd(w_ptr, g_ptr, mom_ptr, n, float(lr), float(momentum), float(weight_decay),
                         float(grad_scale), bool(zero_grad), blocks, _stream_handle(stream))
    launch_counter.add("fed_local_sgd")
