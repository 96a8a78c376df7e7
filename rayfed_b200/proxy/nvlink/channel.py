"""DeviceChannel — the on-box data plane of one party.

Layout of the (symmetric) channel buffer, identical in every party::

    [0      , 256 )  flags[s]   u32  written by *source* s: #messages from s completed
    [256    , 512 )  credits[d] u64  written by *dest* d: bytes of my sends d has consumed
    [512    , 1024)  status u32 | push_done u32 | unpack_done u32      (local use)
    [1024   , 1280)  bflags[s]   u32  like flags, for the broadcast ring of source s
    [1280   , 1536)  bcredits[d] u64  like credits, for my broadcast ring
    [64 KiB , 128K)  signal pad (K5 barrier / K2 aggregate per-CTA flags)
    [128 KiB, ... )  arena: one unicast region of R bytes per source party, then one
                     broadcast region of R bytes per source party (K1b: written once with
                     multimem.st, lands at the same offset in every party)

Send p→d: the sender owns the allocation of ``region[p]`` in d's arena (a byte
ring with monotonically increasing ``head``), so no negotiation is needed — the
deterministic-order property the fed layer already has.  The push
kernel stores the payload into peer memory and release-stores ``flags[p] = k``
for the k-th message.  The receiver's unpack kernel acquires the flag *on its
own stream*, copies/dequantises into a fresh tensor and release-stores
``credits[d] = end_head`` back into the sender's buffer.  The sender only
consults credits when the ring looks full.
"""
from __future__ import annotations

import logging
import threading
import time
from typing import List, Optional, Tuple

from ... import _tracing
from ...ops import kernels as K
from ...parallel.group import PartyGroup
from ...parallel.symmetric import SymmetricHeap

logger = logging.getLogger(__name__)

OFF_FLAGS = 0
OFF_CREDITS = 256
OFF_STATUS = 512
OFF_PUSH_DONE = 516
OFF_UNPACK_DONE = 520
OFF_BFLAGS = 1024  # bflags[s] u32: #broadcast messages from source s completed
OFF_BCREDITS = 1280  # bcredits[d] u64 (in the *sender's* buffer): bytes of my broadcasts d has consumed
OFF_SIGNAL = 64 * 1024
OFF_ARENA = 128 * 1024
SIGNAL_BYTES = 64 * 1024
ALIGN = 256

_WIRE = {None: None, "none": None, "bf16": "bf16", "fp8_e4m3": "fp8_e4m3", "fp8": "fp8_e4m3"}


class DeviceChannel:
    def __init__(self, group: PartyGroup, heap: Optional[SymmetricHeap] = None, region_bytes: int = 256 << 20,
                 wire_dtype: Optional[str] = None, timeout_s: float = 120.0, multicast: bool = True):
        import torch

        self.group = group
        self.heap = heap or SymmetricHeap(group)
        self.region_bytes = int(region_bytes) // ALIGN * ALIGN
        self.wire_dtype = _WIRE[wire_dtype]
        self._barena_off = OFF_ARENA + group.size * self.region_bytes
        total = OFF_ARENA + 2 * group.size * self.region_bytes
        # VMM + NVLS multicast when the box supports it (falls back to cudaIpc collectively)
        self.buf = self.heap.alloc("fed-channel", total, multicast=multicast and group.size > 1)
        self._bhead = 0
        self._bcredit_seen = [0] * group.size
        self._bseq = 0
        self.device_index = group.device_index
        self._dev = f"cuda:{self.device_index}"
        self._send_lock = threading.Lock()
        self._head = [0] * group.size  # bytes allocated towards each dest (monotonic)
        self._credit_seen = [0] * group.size
        self._seq_out = [0] * group.size  # messages pushed towards each dest
        self._seq_in = [0] * group.size  # messages expected from each source
        self._recv_lock = threading.Lock()
        self._push_stream = torch.cuda.Stream(device=self._dev)
        self._recv_stream = torch.cuda.Stream(device=self._dev)
        self._epoch = 0
        # error word in pinned host memory (UVA: same pointer on the device): kernels
        # store to it on a wait timeout, the host reads it with a plain load
        self._status_host = torch.zeros(16, dtype=torch.int32).pin_memory()
        self._timeout_ns = int(timeout_s * 1e9)  # device waits use %globaltimer (ns)
        self.stats = {"tensors_sent": 0, "bytes_sent": 0, "tensors_received": 0, "bytes_received": 0,
                      "credit_stalls": 0, "bcast_tensors_sent": 0, "bcast_bytes_sent": 0}

    # -- addresses ---------------------------------------------------------------
    def _peer(self, party_rank: int, off: int) -> int:
        return self.buf.peer_ptrs[party_rank] + off

    def _local(self, off: int) -> int:
        return self.buf.local_ptr + off

    @property
    def signal_ptr(self) -> int:
        return self._local(OFF_SIGNAL)

    @property
    def signal_peers(self) -> List[int]:
        return [p + OFF_SIGNAL for p in self.buf.peer_ptrs]

    def next_epoch(self, n: int = 2) -> int:
        e = self._epoch
        self._epoch += n
        return e

    # -- send ----------------------------------------------------------------------
    def _alloc(self, dest: int, nbytes: int) -> Tuple[int, int]:
        """Reserve ``nbytes`` in my region of dest's arena → (offset_in_region, end_head)."""
        need = (nbytes + ALIGN - 1) // ALIGN * ALIGN
        if need > self.region_bytes:
            raise ValueError(
                f"tensor of {nbytes} bytes exceeds the device arena region ({self.region_bytes} bytes); "
                "raise cross_silo_comm.device_arena_bytes"
            )
        head = self._head[dest]
        off = head % self.region_bytes
        if off + need > self.region_bytes:  # do not wrap inside a message: skip the tail
            head += self.region_bytes - off
            off = 0
        end = head + need
        deadline = time.time() + 120
        while end - self._credit_seen[dest] > self.region_bytes:
            # ring looks full: refresh the credit word the receiver writes into *my* buffer
            self._credit_seen[dest] = K.cuda_ext().read_u64(self._local(OFF_CREDITS + 8 * dest))
            if end - self._credit_seen[dest] <= self.region_bytes:
                break
            self.stats["credit_stalls"] += 1
            if time.time() > deadline:
                raise TimeoutError("device arena full: the receiver is not consuming")
            time.sleep(0.0002)
        self._head[dest] = end
        return off, end

    def push_tensor(self, dest_rank: int, tensor, ready_event=None, scale: float = 1.0,
                    producer_stream=None) -> dict:
        """Launch K1 for one CUDA tensor; returns the descriptor the host frame carries."""
        import torch

        with torch.cuda.device(self.device_index):  # callers are arbitrary runtime threads
            return self._push_tensor(dest_rank, tensor, ready_event, scale, producer_stream)

    def _push_tensor(self, dest_rank: int, tensor, ready_event=None, scale: float = 1.0,
                     producer_stream=None) -> dict:
        import torch

        t = tensor if tensor.is_contiguous() else tensor.contiguous()
        if t.data_ptr() % 16:
            t = t.clone()
        numel = t.numel()
        wire = self.wire_dtype
        mode = K.MODE_COPY
        payload_bytes = numel * t.element_size()
        scales_bytes = 0
        n_arg = payload_bytes
        if wire == "bf16" and t.dtype == torch.float32:
            mode, payload_bytes, n_arg = K.MODE_F32_TO_BF16, numel * 2, numel
        elif wire == "fp8_e4m3" and t.dtype == torch.bfloat16 and numel % 32 == 0 and numel > 0:
            mode, payload_bytes, n_arg = K.MODE_BF16_TO_FP8, numel, numel
            scales_bytes = numel // 32
        elif scale != 1.0 and t.dtype == torch.bfloat16:
            mode, n_arg = K.MODE_BF16_SCALE, numel
        with self._send_lock:
            pay_pad = (payload_bytes + ALIGN - 1) // ALIGN * ALIGN
            off, end_head = self._alloc(dest_rank, pay_pad + scales_bytes)
            self._seq_out[dest_rank] += 1
            seq = self._seq_out[dest_rank]
            region = OFF_ARENA + self.group.rank * self.region_bytes
            dst = self._peer(dest_rank, region + off)
            dst_scales = dst + pay_pad if scales_bytes else 0
            stream = self._push_stream
            if ready_event is not None:
                stream.wait_event(ready_event)
            else:
                stream.wait_stream(torch.cuda.current_stream(self._dev))
            tl = _tracing._TRACER
            if tl is not None:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record(stream)
            K.push(t.data_ptr(), dst, n_arg, mode=mode, scale=scale, dst_scales_ptr=dst_scales,
                   flag_ptr=self._peer(dest_rank, OFF_FLAGS + 4 * self.group.rank), flag_value=seq,
                   done_counter_ptr=self._local(OFF_PUSH_DONE), stream=stream, device_index=self.device_index)
            if tl is not None:
                ev1 = torch.cuda.Event(enable_timing=True)
                ev1.record(stream)
                tl.device_span("fed_push", ev0, ev1, dest=dest_rank, seq=seq, bytes=payload_bytes + scales_bytes)
            t.record_stream(stream)
        self.stats["tensors_sent"] += 1
        self.stats["bytes_sent"] += payload_bytes + scales_bytes
        return {
            "off": off, "seq": seq, "end": end_head, "mode": mode, "numel": numel, "dtype": str(t.dtype)[6:],
            "shape": tuple(tensor.shape), "pay": payload_bytes, "pad": pay_pad, "scales": scales_bytes,
        }

    # -- broadcast (K1b) -------------------------------------------------------------
    @property
    def can_multicast(self) -> bool:
        return bool(self.buf.mc_ptr) and self.group.size > 1

    def _alloc_bcast(self, nbytes: int):
        need = (nbytes + ALIGN - 1) // ALIGN * ALIGN
        if need > self.region_bytes:
            raise ValueError("tensor exceeds the broadcast arena region")
        head = self._bhead
        off = head % self.region_bytes
        if off + need > self.region_bytes:
            head += self.region_bytes - off
            off = 0
        end = head + need
        deadline = time.time() + 120
        peers = [d for d in range(self.group.size) if d != self.group.rank]
        while end - min(self._bcredit_seen[d] for d in peers) > self.region_bytes:
            for d in peers:  # every receiver must have consumed before the ring wraps
                self._bcredit_seen[d] = K.cuda_ext().read_u64(self._local(OFF_BCREDITS + 8 * d))
            if end - min(self._bcredit_seen[d] for d in peers) <= self.region_bytes:
                break
            self.stats["credit_stalls"] += 1
            if time.time() > deadline:
                raise TimeoutError("broadcast arena full: a receiver is not consuming")
            time.sleep(0.0002)
        self._bhead = end
        return off, end

    def can_bcast_tensor(self, tensor) -> bool:
        nbytes = tensor.numel() * tensor.element_size()
        return (self.can_multicast and self.wire_dtype is None and tensor.is_contiguous() and nbytes > 0
                and nbytes % 16 == 0 and tensor.data_ptr() % 16 == 0 and nbytes <= self.region_bytes)

    def push_tensor_bcast(self, tensor, ready_event=None) -> Optional[dict]:
        """K1b: ONE multimem.st kernel puts the tensor into every party's broadcast region.
        Returns the descriptor, or None when this tensor has to go unicast."""
        import torch

        if not self.can_multicast or self.wire_dtype is not None:
            return None
        with torch.cuda.device(self.device_index):
            t = tensor if tensor.is_contiguous() else tensor.contiguous()
            nbytes = t.numel() * t.element_size()
            if nbytes % 16 or t.data_ptr() % 16 or nbytes == 0:
                return None
            with self._send_lock:
                off, end_head = self._alloc_bcast(nbytes)
                self._bseq += 1
                seq = self._bseq
                slot = self._barena_off + self.group.rank * self.region_bytes + off
                stream = self._push_stream
                if ready_event is not None:
                    stream.wait_event(ready_event)
                else:
                    stream.wait_stream(torch.cuda.current_stream(self._dev))
                flags = [self._peer(d, OFF_BFLAGS + 4 * self.group.rank) for d in range(self.group.size)
                         if d != self.group.rank]
                K.push_mc(t.data_ptr(), self.buf.mc_ptr + slot, nbytes, flag_ptrs=flags, flag_value=seq,
                          done_counter_ptr=self._local(OFF_PUSH_DONE), stream=stream,
                          device_index=self.device_index)
                t.record_stream(stream)
            self.stats["bcast_tensors_sent"] += 1
            self.stats["bcast_bytes_sent"] += nbytes
            return {
                "off": off, "seq": seq, "end": end_head, "mode": K.MODE_COPY, "numel": t.numel(),
                "dtype": str(t.dtype)[6:], "shape": tuple(tensor.shape), "pay": nbytes,
                "pad": (nbytes + ALIGN - 1) // ALIGN * ALIGN, "scales": 0, "bcast": True,
            }

    def record_push_event(self):
        """Event after everything launched so far on the push stream."""
        import torch

        with torch.cuda.device(self.device_index):
            ev = torch.cuda.Event()
            ev.record(self._push_stream)
        return ev

    # -- receive -----------------------------------------------------------------------
    def unpack_tensor(self, src_rank: int, desc: dict):
        """Launch K4+unpack for one descriptor → (tensor, event)."""
        import torch

        with torch.cuda.device(self.device_index):
            return self._unpack_tensor(src_rank, desc)

    def _unpack_tensor(self, src_rank: int, desc: dict):
        import torch

        dtype = getattr(torch, desc["dtype"])
        out = torch.empty(desc["shape"], dtype=dtype, device=self._dev)
        bcast = bool(desc.get("bcast"))
        region = (self._barena_off if bcast else OFF_ARENA) + src_rank * self.region_bytes
        src = self._local(region + desc["off"])
        flag_off = (OFF_BFLAGS if bcast else OFF_FLAGS) + 4 * src_rank
        credit_off = (OFF_BCREDITS if bcast else OFF_CREDITS) + 8 * self.group.rank
        mode = desc["mode"]
        if mode == K.MODE_F32_TO_BF16:
            un_mode, n_arg, scales = K.MODE_BF16_TO_F32, desc["numel"], 0
        elif mode == K.MODE_BF16_TO_FP8:
            un_mode, n_arg, scales = K.MODE_FP8_TO_BF16, desc["numel"], src + desc["pad"]
        else:
            un_mode, n_arg, scales = K.MODE_COPY, desc["numel"] * out.element_size(), 0
        with self._recv_lock:
            stream = self._recv_stream
            # the fresh tensor's memory may still be in use by earlier work on the caller's stream
            stream.wait_stream(torch.cuda.current_stream(self._dev))
            K.unpack(src, out.data_ptr(), n_arg, mode=un_mode, scales_ptr=scales,
                     flag_ptr=self._local(flag_off), flag_value=desc["seq"],
                     done_counter_ptr=self._local(OFF_UNPACK_DONE),
                     credit_ptr=self._peer(src_rank, credit_off), credit_value=desc["end"],
                     status_ptr=self._status_host.data_ptr(), timeout_ns=self._timeout_ns, stream=stream,
                     device_index=self.device_index)
            ev = torch.cuda.Event()
            ev.record(stream)
            out.record_stream(stream)
        self.stats["tensors_received"] += 1
        self.stats["bytes_received"] += desc["pay"] + desc["scales"]
        return out, ev

    # -- misc --------------------------------------------------------------------------
    def device_barrier(self, stream=None):
        """K5 on the given (default: current) stream."""
        K.cuda_ext().set_device(self.device_index)
        K.barrier(n_parties=self.group.size, my_rank=self.group.rank, sig_ptr=self.signal_ptr,
                  sig_peers=self.signal_peers, epoch=self.next_epoch(1), status_ptr=self._status_host.data_ptr(),
                  timeout_ns=self._timeout_ns, stream=stream)

    def check_status(self) -> int:
        """0 = healthy; 1 = a device-side receive wait timed out; 2 = a party barrier did."""
        return int(self._status_host[0])

    def raise_if_failed(self):
        st = self.check_status()
        if st:
            from ...exceptions import FedError

            raise FedError(
                f"NVLink device channel failure (status {st}): a device-side wait for a peer timed out after "
                f"{self._timeout_ns / 1e9:.0f}s — the peer party is gone or its producer never launched."
            )

    def synchronize(self):
        self._push_stream.synchronize()
        self._recv_stream.synchronize()
