"""DeviceChannel — the on-box data plane of one party (K1 push / K4 wait / K1b multicast).

The reference moves every cross-party object through ``D2H → cloudpickle → protobuf → HTTP/2 →
event table → unpickle → H2D`` (fed/proxy/grpc/grpc_proxy.py:193-220 sender, :272-342
receiver).  Here the tensor bytes of an object are stored by ONE kernel straight into the
consumer party's arena in peer HBM, preceded by a fixed-layout 64-byte descriptor and followed by
a release flag; the consumer orders its stream on that flag with a *stream memory operation*
(``cuStreamWaitValue32`` — no SM spins, nothing blocks on the host) and hands the bytes out as
**zero-copy views of the arena**.  No host thread is on the path between "producer kernel
finished" and "consumer kernel may start".

Layout of the (symmetric) channel buffer, identical in every party::

    [0      , 64  )  flags[s]       u32  written by *source* s: #messages s→me completed
    [64     , 128 )  bflags[s]      u32  same, for broadcast messages of source s
    [256    , 384 )  freed_cnt[d]   u64  written by *dest* d: #entries d appended to freed_ring[d]
    [384    , 512 )  bfreed_cnt[d]  u64  same for my broadcast blocks
    [512    , 1024)  local words: push_done, unpack_done
    [4 KiB  , 36 K)  freed_ring[d][256]  u64  ids of my blocks that dest d has released
    [36 KiB , 68 K)  bfreed_ring[d][256] u64
    [128 KiB, 192K)  signal pad (K5 barrier / K2 aggregate per-CTA flags)
    [1 MiB  , ... )  arena: one unicast region of R bytes per source party, then one broadcast
                     region of R bytes per source party (K1b: written once with multimem.st,
                     lands at the same offset in every party)

Send p→d.  The *sender* owns the allocation of ``region[p]`` in d's arena (a free-list allocator:
blocks are released out of order, because a zero-copy view lives as long as the user keeps the
tensor).  A message is one block ``[256-byte header | tensor | tensor | ...]``; the push kernel
stores the payload, then the descriptor into the header, then release-stores ``flags[p] = seq``.
Its metadata (block id, offsets, shapes) goes out on the host transport *at launch time* — the
host never waits for the push kernel.

Receive.  A per-source thread takes the metadata frames in order and enqueues, on that source's
receive stream: ``cuStreamWaitValue32(flags[src] >= seq)``, a one-warp descriptor check, and
either nothing more (zero-copy views) or the unpack kernels for wire formats / small / chunked
tensors.  The consumer's stream waits for the event recorded behind that.

Flow control.  When the last view of a block dies (or its copy-out is enqueued) the credit
stream is ordered after the consumers and a one-thread kernel appends the block id to the
``freed_ring`` in the *sender's* control block.  The sender reads that ring only when an
allocation does not fit.
"""
from __future__ import annotations

import collections
import logging
import os
import struct
import threading
import time
import weakref
from typing import Dict, List, Optional, Tuple

from ... import _tracing
from ...ops import kernels as K
from ...parallel.group import PartyGroup
from ...parallel.symmetric import SymmetricHeap

logger = logging.getLogger(__name__)

MAX_PARTIES = 16
OFF_FLAGS = 0
OFF_BFLAGS = 64
OFF_FREED_CNT = 256
OFF_BFREED_CNT = 384
OFF_PUSH_DONE = 512
OFF_UNPACK_DONE = 520
OFF_FREED_RING = 4 * 1024
OFF_BFREED_RING = 36 * 1024
FREED_RING = 256  # entries per destination
OFF_SIGNAL = 128 * 1024
SIGNAL_BYTES = 64 * 1024
OFF_ARENA = 1 << 20
ALIGN = 256
HEADER = 256  # block header: 64-byte descriptor + padding (payload stays 256-byte aligned)
MAX_SEGS = 48

_WIRE = {None: None, "none": None, "bf16": "bf16", "fp8_e4m3": "fp8_e4m3", "fp8": "fp8_e4m3"}
_DTYPE_CODE = {"uint8": 1, "int8": 2, "int16": 3, "int32": 4, "int64": 5, "float16": 6, "bfloat16": 7,
               "float32": 8, "float64": 9, "bool": 10, "float8_e4m3fn": 11, "float8_e5m2": 12}
DESC_MAGIC = 0x46454450
_DESC = struct.Struct("<IIIIIIQQBBBBI4I")
assert _DESC.size == 64


def fnv1a32(s: str) -> int:
    h = 0x811C9DC5
    for b in s.encode():
        h = ((h ^ b) * 0x01000193) & 0xFFFFFFFF
    return h


def _key_words(upstream, downstream) -> Tuple[int, int, int]:
    """(up_task, up_idx, down) of a fed message key; ids that are not ``task#idx`` / int are hashed."""
    up, idx = str(upstream), 0
    if "#" in up:
        a, _, b = up.partition("#")
        if a.isdigit() and b.isdigit():
            up, idx = a, int(b)
    up_w = int(up) & 0xFFFFFFFF if up.isdigit() else fnv1a32(up)
    down = str(downstream)
    down_w = int(down) & 0xFFFFFFFF if down.isdigit() else fnv1a32(down)
    return up_w, idx & 0xFFFFFFFF, down_w


def pack_desc(*, seq, job_hash, key, offset, nbytes, dtype="uint8", mode=0, ndim=0, n_tensors=1, shape=()) -> bytes:
    up_task, up_idx, down = key
    dims = [int(d) & 0xFFFFFFFF for d in tuple(shape)[:4]]
    dims += [0] * (4 - len(dims))
    return _DESC.pack(DESC_MAGIC, seq & 0xFFFFFFFF, job_hash, up_task, up_idx, down, offset, nbytes,
                      _DTYPE_CODE.get(dtype, 0), mode & 0xFF, min(int(ndim), 255), 1, n_tensors, *dims)


def unpack_desc(raw: bytes) -> dict:
    f = _DESC.unpack(raw)
    return {"magic": f[0], "seq": f[1], "job_hash": f[2], "up_task": f[3], "up_idx": f[4], "down": f[5],
            "offset": f[6], "nbytes": f[7], "dtype": f[8], "mode": f[9], "ndim": f[10], "status": f[11],
            "n_tensors": f[12], "shape": list(f[13:17])}


def _pad(n: int) -> int:
    return (n + ALIGN - 1) // ALIGN * ALIGN


class RegionAllocator:
    """Free-list allocator over one region of a peer's arena (sender side).  Next-fit from a
    moving cursor, so traffic whose blocks are released in order behaves like a ring; blocks that
    stay alive (a consumer keeps a zero-copy view) are simply skipped."""

    def __init__(self, size: int):
        self.size = size
        self._free: List[List[int]] = [[0, size]]  # sorted [off, len]
        self._blocks: Dict[int, Tuple[int, int]] = {}
        self._next_bid = 1
        self._cursor = 0

    @property
    def outstanding(self) -> int:
        return len(self._blocks)

    @property
    def free_bytes(self) -> int:
        return sum(f[1] for f in self._free)

    def try_alloc(self, nbytes: int) -> Optional[Tuple[int, int]]:
        need = _pad(nbytes)
        n = len(self._free)
        if n == 0:
            return None
        start = 0
        for i, f in enumerate(self._free):
            if f[0] + f[1] > self._cursor:
                start = i
                break
        for j in range(n):
            i = (start + j) % n
            off, ln = self._free[i]
            if j == 0 and off < self._cursor < off + ln:
                # the cursor sits inside this hole: use the part behind it first
                tail = off + ln - self._cursor
                if tail >= need:
                    got = self._cursor
                    pieces = [[off, got - off], [got + need, tail - need]]
                    self._free[i:i + 1] = [p for p in pieces if p[1] > 0]
                    return self._take(got, need)
                continue
            if ln >= need:
                if ln == need:
                    del self._free[i]
                else:
                    self._free[i] = [off + need, ln - need]
                return self._take(off, need)
        # last resort: the head part of the hole the cursor sits in
        off, ln = self._free[start]
        if ln >= need:
            if ln == need:
                del self._free[start]
            else:
                self._free[start] = [off + need, ln - need]
            return self._take(off, need)
        return None

    def _take(self, off: int, need: int) -> Tuple[int, int]:
        bid = self._next_bid
        self._next_bid += 1
        self._blocks[bid] = (off, need)
        self._cursor = (off + need) % self.size
        return bid, off

    def free(self, bid: int) -> bool:
        blk = self._blocks.pop(bid, None)
        if blk is None:
            return False
        off, ln = blk
        fr = self._free
        lo, hi = 0, len(fr)
        while lo < hi:
            mid = (lo + hi) // 2
            if fr[mid][0] < off:
                lo = mid + 1
            else:
                hi = mid
        fr.insert(lo, [off, ln])
        if lo + 1 < len(fr) and fr[lo][0] + fr[lo][1] == fr[lo + 1][0]:
            fr[lo][1] += fr[lo + 1][1]
            del fr[lo + 1]
        if lo > 0 and fr[lo - 1][0] + fr[lo - 1][1] == fr[lo][0]:
            fr[lo - 1][1] += fr[lo][1]
            del fr[lo]
        return True


class _ArenaBlock:
    """``__cuda_array_interface__`` over one received block.  torch keeps this object alive for
    as long as any view of the block exists; its death is the block's release."""

    __slots__ = ("__cuda_array_interface__", "_on_free", "__weakref__")

    def __init__(self, ptr: int, nbytes: int, on_free):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}
        self._on_free = on_free

    def __del__(self):
        cb, self._on_free = self._on_free, None
        if cb is not None:
            try:
                cb()
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass


class OwnedEvent:
    """A raw CUDA event (timing disabled) destroyed with its last Python reference.  Consumers
    hold it through ``DeviceValue.owner`` / the Ref's device event; destroying an event that
    still has waits pending is legal — the driver defers the release."""

    __slots__ = ("handle", "_ext")

    def __init__(self, ext):
        self._ext = ext
        self.handle = ext.event_create(False)

    def __del__(self):
        try:
            self._ext.event_destroy(self.handle)
        except Exception:  # noqa: BLE001  (context already gone at interpreter exit)
            pass


class DeviceChannel:
    def __init__(self, group: PartyGroup, heap: Optional[SymmetricHeap] = None, region_bytes: int = 256 << 20,
                 wire_dtype: Optional[str] = None, timeout_s: float = 120.0, multicast: bool = True,
                 job_name: str = "", zero_copy: bool = True, zero_copy_min_bytes: int = 64 * 1024,
                 consumer_streams=None, wait_timeout_s: Optional[float] = None):
        import torch

        if group.size > MAX_PARTIES:
            raise ValueError(f"the device channel supports at most {MAX_PARTIES} parties on a box (got {group.size})")
        self.group = group
        self.heap = heap or SymmetricHeap(group)
        self.region_bytes = int(region_bytes) // ALIGN * ALIGN
        self.wire_dtype = _WIRE[wire_dtype]
        self.job_hash = fnv1a32(job_name or "")
        self.zero_copy = bool(zero_copy)
        self.zero_copy_min_bytes = int(zero_copy_min_bytes)
        self._consumer_streams = consumer_streams or (lambda: [])
        self._barena_off = OFF_ARENA + group.size * self.region_bytes
        total = OFF_ARENA + 2 * group.size * self.region_bytes
        # VMM + NVLS multicast when the box supports it (falls back to cudaIpc collectively)
        self.buf = self.heap.alloc("fed-channel", total, multicast=multicast and group.size > 1)
        self.device_index = group.device_index
        self._dev = f"cuda:{self.device_index}"
        self._ext = K.cuda_ext()
        with torch.cuda.device(self.device_index):
            self._ext.preload_kernels()  # no first launch (= lazy module load) under a parked wait
            # LAZY module loading: the first launch of any kernel blocks its host thread while a stream
            # wait is parked (measured), which can deadlock two parties in their first round.  Then
            # receives are gated on the host: the flag is polled (tiny D2H reads on a private stream)
            # and the — by then satisfied — stream wait is enqueued afterwards; nothing ever parks.
            # With CUDA_MODULE_LOADING=EAGER the stream waits are enqueued at once (no host in the path).
            mode = os.environ.get("RAYFED_B200_DEVICE_WAIT", "auto")  # auto | stream | host
            try:
                lazy = self._ext.module_loading_mode() != "eager"
            except Exception:  # noqa: BLE001  (old driver without the query)
                lazy = True
            self._host_gate = (mode == "host") or (mode == "auto" and lazy)
            if self._host_gate:
                logger.info("device channel: CUDA module loading is lazy -> receives are host-gated "
                            "(set CUDA_MODULE_LOADING=EAGER before the first CUDA call for parked stream waits)")
        self._timeout_s = float(timeout_s)
        self._timeout_ns = int(timeout_s * 1e9)  # device waits use %globaltimer (ns)
        # A parked receive legitimately lasts as long as the producer's device work (the metadata
        # is announced at launch time), so its watchdog is far more patient than an RPC timeout.
        self._wait_timeout_s = float(wait_timeout_s) if wait_timeout_s else max(10.0 * self._timeout_s, 600.0)
        self._tstream_cache = {}
        n = group.size
        # -- send side
        self._send_lock = threading.Lock()
        self._alloc = [RegionAllocator(self.region_bytes) for _ in range(n)]
        self._balloc = RegionAllocator(self.region_bytes)
        self._bfreed_by: Dict[int, set] = {}
        self._seq_out = [0] * n
        self._bseq = 0
        self._freed_seen = [0] * n
        self._bfreed_seen = [0] * n
        self._push_stream = torch.cuda.Stream(device=self._dev)
        # -- receive side: one raw (non-pool) stream per source, so a parked wait on one source
        # never delays deliveries from another
        self._recv_lock = threading.Lock()
        self._gate_lock = threading.Lock()
        with torch.cuda.device(self.device_index):
            self._recv_streams = [self._ext.stream_create(0) if s != group.rank else 0 for s in range(n)]
            self._recv_tstreams = [torch.cuda.ExternalStream(h, device=self._dev) if h else None
                                   for h in self._recv_streams]
            self._credit_stream = self._ext.stream_create(0)
            self._poll_stream = self._ext.stream_create(0)
            self._gate_stream = self._ext.stream_create(0)
        self._view_bytes = [0] * n  # bytes currently handed out as zero-copy views, per source
        self._freed_idx = [0] * n  # entries I appended to freed_ring[me] in source s's block
        self._bfreed_idx = [0] * n
        self._credit_q = collections.deque()
        self._credit_wake = threading.Event()
        self._pending_waits = collections.deque()  # (event, t_enqueued, src, seq, bcast)
        self._stopping = False
        self._credit_thread = threading.Thread(target=self._credit_loop, name="fed-credit", daemon=True)
        self._credit_thread.start()
        self._epoch = 0
        # error word in pinned host memory (UVA: same pointer on the device): kernels
        # store to it on a wait timeout / descriptor mismatch, the host reads it with a plain load
        self._status_host = torch.zeros(16, dtype=torch.int32).pin_memory()
        self.push_times = None  # bench: list of (bytes, ev0, ev1) CUDA-event pairs around each push message
        self.stats = {"tensors_sent": 0, "bytes_sent": 0, "messages_sent": 0, "tensors_received": 0,
                      "bytes_received": 0, "messages_received": 0, "zero_copy_tensors": 0, "copied_tensors": 0,
                      "credit_stalls": 0, "bcast_tensors_sent": 0, "bcast_bytes_sent": 0, "blocks_released": 0,
                      "forced_releases": 0, "wait_mode": "host" if self._host_gate else "stream"}

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

    @property
    def can_multicast(self) -> bool:
        return bool(self.buf.mc_ptr) and self.group.size > 1

    @property
    def max_message_bytes(self) -> int:
        """Largest block: bigger tensors are chunked through the arena (pipelined with the
        consumer's copy-out) instead of spilling to the host path."""
        return max(self.region_bytes // 4, 1 << 20)

    # -- send ----------------------------------------------------------------------
    def _poll_freed(self, dest: int, bcast: bool = False) -> int:
        """Read the freed ring dest wrote into MY control block; returns #blocks recycled."""
        ext = self._ext
        cnt_off = (OFF_BFREED_CNT if bcast else OFF_FREED_CNT) + 8 * dest
        ring_off = (OFF_BFREED_RING if bcast else OFF_FREED_RING) + dest * FREED_RING * 8
        seen = self._bfreed_seen if bcast else self._freed_seen
        cnt = ext.read_u64s(self._local(cnt_off), 1, self._poll_stream)[0]
        new = cnt - seen[dest]
        if new <= 0:
            return 0
        if new > FREED_RING:
            raise RuntimeError("freed ring overrun: flow-control invariant violated")
        ring = ext.read_u64s(self._local(ring_off), FREED_RING, self._poll_stream)
        n = 0
        for i in range(seen[dest], cnt):
            bid = ring[i % FREED_RING]
            if bcast:
                got = self._bfreed_by.setdefault(bid, set())
                got.add(dest)
                if len(got) >= self.group.size - 1:
                    self._bfreed_by.pop(bid, None)
                    n += 1 if self._balloc.free(bid) else 0
            else:
                n += 1 if self._alloc[dest].free(bid) else 0
        seen[dest] = cnt
        return n

    def _alloc_block(self, dest: Optional[int], nbytes: int) -> Tuple[int, int]:
        """(block id, offset in my region of dest's arena); ``dest=None``: my broadcast region.
        Waits — with the configured timeout — while the consumers hold too much of the region."""
        bcast = dest is None
        alloc = self._balloc if bcast else self._alloc[dest]
        if _pad(nbytes) > self.region_bytes:
            raise ValueError(f"block of {nbytes} bytes exceeds the arena region ({self.region_bytes} bytes)")
        peers = [d for d in range(self.group.size) if d != self.group.rank] if bcast else [dest]
        deadline = None
        pause = 20e-6
        while True:
            if alloc.outstanding < FREED_RING - 64:
                got = alloc.try_alloc(nbytes)
                if got is not None:
                    return got
            recycled = sum(self._poll_freed(d, bcast) for d in peers)
            if recycled:
                continue
            self.stats["credit_stalls"] += 1
            now = time.monotonic()
            if deadline is None:
                deadline = now + self._timeout_s
            elif now > deadline:
                raise TimeoutError(
                    f"device arena full for {self._timeout_s:.0f}s: the consumer party is not releasing blocks "
                    f"({alloc.outstanding} outstanding, {alloc.free_bytes} bytes free, need {nbytes})")
            time.sleep(pause)
            pause = min(pause * 2, 2e-3)

    def _plan_tensor(self, t):
        """(contiguous tensor, wire mode, payload bytes, scales bytes, n_arg per element unit)."""
        import torch

        t = t if t.is_contiguous() else t.contiguous()
        if t.data_ptr() % 16:
            t = t.clone()
        numel = t.numel()
        wire = self.wire_dtype
        if wire == "bf16" and t.dtype == torch.float32:
            return t, K.MODE_F32_TO_BF16, numel * 2, 0
        if wire == "fp8_e4m3" and t.dtype == torch.bfloat16 and numel % 32 == 0 and numel > 0:
            return t, K.MODE_BF16_TO_FP8, numel, numel // 32
        return t, K.MODE_COPY, numel * t.element_size(), 0

    def plan_messages(self, tensors) -> List[List[dict]]:
        """Pack an object's tensors into messages (= arena blocks).  A part is a whole tensor or —
        for tensors larger than a block — an element range of one."""
        cap = self.max_message_bytes - HEADER
        messages, cur, cur_bytes = [], [], 0

        def flush():
            nonlocal cur, cur_bytes
            if cur:
                messages.append(cur)
                cur, cur_bytes = [], 0

        for ti, t in enumerate(tensors):
            tt, mode, pay, scales = self._plan_tensor(t)
            size = _pad(pay) + _pad(scales)
            if size <= cap:
                if cur_bytes + size > cap or len(cur) >= MAX_SEGS:
                    flush()
                cur.append({"t": ti, "e0": 0, "ne": tt.numel(), "mode": mode, "pay": pay, "scales": scales,
                            "whole": True, "tensor": tt})
                cur_bytes += size
                continue
            flush()
            # chunk by elements; multiples of 256 keep 16-byte alignment and fp8 scale blocks whole
            esz_wire = pay / max(tt.numel(), 1)
            per = max(256, int(cap / (esz_wire + (1 / 32 if scales else 0))) // 256 * 256)
            e0 = 0
            while e0 < tt.numel():
                ne = min(per, tt.numel() - e0)
                p = ne * tt.element_size() if mode == K.MODE_COPY else (ne * 2 if mode == K.MODE_F32_TO_BF16 else ne)
                messages.append([{"t": ti, "e0": e0, "ne": ne, "mode": mode, "pay": p,
                                  "scales": ne // 32 if scales else 0, "whole": False, "tensor": tt}])
                e0 += ne
        flush()
        return messages

    @staticmethod
    def layout_parts(parts: List[dict]) -> int:
        """Place the parts behind the block header; returns the block size."""
        toff = HEADER
        for p in parts:
            p["toff"] = toff
            p["pad"] = _pad(p["pay"])
            toff += p["pad"] + _pad(p["scales"])
        return toff

    @staticmethod
    def part_meta(p: dict) -> dict:
        """What the host frame says about one part (the receiver validates it, `_validate_meta`)."""
        t = p.pop("tensor")
        return {"t": p["t"], "e0": p["e0"], "ne": p["ne"], "mode": p["mode"], "pay": p["pay"],
                "pad": p["pad"], "scales": p["scales"], "toff": p["toff"], "whole": p["whole"],
                "numel": t.numel(), "dtype": str(t.dtype)[6:], "shape": list(t.shape)}

    def can_bcast(self, tensors) -> bool:
        """K1b applies to plain byte copies whose sizes the 16-byte multimem stores cover exactly."""
        if not self.can_multicast or self.wire_dtype is not None:
            return False
        for t in tensors:
            nbytes = t.numel() * t.element_size()
            if nbytes == 0 or nbytes % 16:
                return False
        return True

    def push_message(self, dest_rank: Optional[int], parts: List[dict], key, ready_event: int = 0,
                     producer_stream: int = 0) -> dict:
        """Launch K1 (or K1b when ``dest_rank`` is None) for one message; returns the metadata the
        host frame carries.  Never waits for the device."""
        import torch

        with torch.cuda.device(self.device_index):  # callers are arbitrary runtime threads
            return self._push_message(dest_rank, parts, key, ready_event, producer_stream)

    def _push_message(self, dest_rank, parts, key, ready_event, producer_stream):
        import torch

        bcast = dest_rank is None
        ext = self._ext
        metas = []
        total = self.layout_parts(parts)
        with self._send_lock:
            bid, off = self._alloc_block(dest_rank, total)
            if bcast:
                self._bseq += 1
                seq = self._bseq
                base = self.buf.mc_ptr + self._barena_off + self.group.rank * self.region_bytes + off
                flags = [self._peer(d, OFF_BFLAGS + 4 * self.group.rank) for d in range(self.group.size)
                         if d != self.group.rank]
            else:
                self._seq_out[dest_rank] += 1
                seq = self._seq_out[dest_rank]
                base = self._peer(dest_rank, OFF_ARENA + self.group.rank * self.region_bytes + off)
                flag = self._peer(dest_rank, OFF_FLAGS + 4 * self.group.rank)
            first = parts[0]["tensor"]
            desc = pack_desc(seq=seq, job_hash=self.job_hash, key=_key_words(*key), offset=off, nbytes=total,
                             dtype=str(first.dtype)[6:], mode=parts[0]["mode"], ndim=first.dim(),
                             n_tensors=len(parts), shape=first.shape)
            stream = self._push_stream
            sh = stream.cuda_stream
            if ready_event:
                if producer_stream != sh:
                    ext.stream_wait_event(sh, ready_event)
            else:
                stream.wait_stream(torch.cuda.current_stream(self._dev))
            tl = _tracing._TRACER
            timed = tl is not None or self.push_times is not None
            if timed:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record(stream)
            done = self._local(OFF_PUSH_DONE)
            nbytes_all = 0
            if bcast:
                for i, p in enumerate(parts):
                    last = i == len(parts) - 1
                    t = p["tensor"]
                    K.push_mc(t.data_ptr() + p["e0"] * t.element_size(), base + p["toff"], p["pay"],
                              flag_ptrs=flags if last else [], flag_value=seq, done_counter_ptr=done, stream=stream,
                              device_index=self.device_index, desc_dst=base if last else 0,
                              desc=desc if last else b"")
                    nbytes_all += p["pay"]
            elif all(p["mode"] == K.MODE_COPY for p in parts) and len(parts) > 1:
                segs = []
                for p in parts:
                    t = p["tensor"]
                    segs.append((t.data_ptr() + p["e0"] * t.element_size(), p["toff"], p["pay"]))
                    nbytes_all += p["pay"]
                K.push_multi(segs, base, flag_ptr=flag, flag_value=seq, done_counter_ptr=done, stream=stream,
                             device_index=self.device_index, desc_dst=base, desc=desc)
            else:
                for i, p in enumerate(parts):
                    last = i == len(parts) - 1
                    t = p["tensor"]
                    mode = p["mode"]
                    src = t.data_ptr() + p["e0"] * t.element_size()
                    n_arg = p["pay"] if mode == K.MODE_COPY else p["ne"]
                    K.push(src, base + p["toff"], n_arg, mode=mode,
                           dst_scales_ptr=base + p["toff"] + p["pad"] if p["scales"] else 0,
                           flag_ptr=flag if last else 0, flag_value=seq, done_counter_ptr=done, stream=stream,
                           device_index=self.device_index, desc_dst=base if last else 0, desc=desc if last else b"")
                    nbytes_all += p["pay"] + p["scales"]
            if timed:
                ev1 = torch.cuda.Event(enable_timing=True)
                ev1.record(stream)
                if tl is not None:
                    tl.device_span("fed_push_mc" if bcast else "fed_push", ev0, ev1, dest=dest_rank, seq=seq,
                                   bytes=nbytes_all)
                if self.push_times is not None:
                    self.push_times.append((nbytes_all, ev0, ev1))
            for p in parts:
                p["tensor"].record_stream(stream)
        st = self.stats
        if bcast:
            st["bcast_tensors_sent"] += len(parts)
            st["bcast_bytes_sent"] += nbytes_all
        else:
            st["tensors_sent"] += len(parts)
            st["bytes_sent"] += nbytes_all
        st["messages_sent"] += 1
        for p in parts:
            metas.append(self.part_meta(p))
        return {"bid": bid, "off": off, "seq": seq, "total": total, "bcast": bcast, "parts": metas,
                "desc": desc.hex()}

    def record_push_event(self):
        """Event after everything launched so far on the push stream (tests, shutdown)."""
        import torch

        with torch.cuda.device(self.device_index):
            ev = torch.cuda.Event()
            ev.record(self._push_stream)
        return ev

    # -- receive -----------------------------------------------------------------------
    def receive_message(self, src_rank: int, meta: dict, assembling: dict):
        """Order the source's receive stream after the message's flag and materialise its parts.
        ``assembling`` maps tensor index → output tensor for tensors that arrive in chunks.
        Returns ``[(tensor index, tensor or None)]`` — ``None`` for a chunk that is not the last."""
        import torch

        with torch.cuda.device(self.device_index):
            return self._receive_message(src_rank, meta, assembling)

    def _receive_message(self, src_rank, meta, assembling):
        import torch

        ext = self._ext
        bcast = bool(meta.get("bcast"))
        seq, bid, off, total = int(meta["seq"]), int(meta["bid"]), int(meta["off"]), int(meta["total"])
        self._validate_meta(meta, off, total)
        region = (self._barena_off if bcast else OFF_ARENA) + src_rank * self.region_bytes
        block = self._local(region + off)
        flag = self._local((OFF_BFLAGS if bcast else OFF_FLAGS) + 4 * src_rank)
        parts = meta["parts"]
        sh = self._recv_streams[src_rank]
        ts = self._recv_tstreams[src_rank]
        if self._host_gate:
            self._gate(flag, seq, src_rank)
        with self._recv_lock:
            # K4: the stream — not a thread, not an SM — waits for the producer's release flag
            ext.stream_wait_value32(sh, flag, seq & 0xFFFFFFFF)
            self.stats["stream_waits"] = self.stats.get("stream_waits", 0) + 1
            ext.desc_check(block, bytes.fromhex(meta["desc"]), self._status_host.data_ptr(), sh)
            K.launch_counter.add("fed_desc_check")
            zero_copy = (self.zero_copy and total >= self.zero_copy_min_bytes
                         and all(p["whole"] and p["mode"] == K.MODE_COPY for p in parts)
                         and self._view_bytes[src_rank] + total <= self.region_bytes // 2)
            out = []
            if zero_copy:
                self._view_bytes[src_rank] += total
                chan = weakref.ref(self)

                def on_free(src=src_rank, bid=bid, total=total, bcast=bcast):
                    ch = chan()
                    if ch is not None:
                        ch._credit_q.append((src, bid, total, bcast, None))
                        ch._credit_wake.set()

                raw = torch.as_tensor(_ArenaBlock(block, total, on_free), device=self._dev)
                for p in parts:
                    dtype = getattr(torch, p["dtype"])
                    v = raw[p["toff"]: p["toff"] + p["pay"]].view(dtype).view(p["shape"])
                    out.append((p["t"], v))
                self.stats["zero_copy_tensors"] += len(parts)
            else:
                with torch.cuda.stream(ts):
                    for p in parts:
                        dtype = getattr(torch, p["dtype"])
                        if p["whole"]:
                            dst = torch.empty(p["shape"], dtype=dtype, device=self._dev)
                            dst_ptr, final = dst.data_ptr(), dst
                        else:
                            full = assembling.get(p["t"])
                            if full is None:
                                full = torch.empty(p["shape"], dtype=dtype, device=self._dev)
                                assembling[p["t"]] = full
                            dst_ptr = full.data_ptr() + p["e0"] * full.element_size()
                            final = full if p["e0"] + p["ne"] >= p["numel"] else None
                            if final is not None:
                                assembling.pop(p["t"], None)
                        mode = p["mode"]
                        src = block + p["toff"]
                        if mode == K.MODE_F32_TO_BF16:
                            un_mode, n_arg, scales = K.MODE_BF16_TO_F32, p["ne"], 0
                        elif mode == K.MODE_BF16_TO_FP8:
                            un_mode, n_arg, scales = K.MODE_FP8_TO_BF16, p["ne"], src + p["pad"]
                        else:
                            un_mode, n_arg, scales = K.MODE_COPY, p["pay"], 0
                        if final is not None:
                            self._mark_consumers(final)
                        K.unpack(src, dst_ptr, n_arg, mode=un_mode, scales_ptr=scales, flag_ptr=0, flag_value=0,
                                 done_counter_ptr=self._local(OFF_UNPACK_DONE), status_ptr=self._status_host.data_ptr(),
                                 timeout_ns=self._timeout_ns, stream=sh, device_index=self.device_index)
                        out.append((p["t"], final))
                self.stats["copied_tensors"] += len(parts)
            ev = OwnedEvent(ext)
            ext.event_record(ev.handle, sh)
            if not zero_copy:
                # the block is free as soon as the copy-outs (ordered before `ev`) are done
                self._credit_q.append((src_rank, bid, 0, bcast, ev))
                self._credit_wake.set()
            self._pending_waits.append((ev, time.monotonic(), src_rank, seq, bcast))
        self.stats["tensors_received"] += len(parts)
        self.stats["messages_received"] += 1
        self.stats["bytes_received"] += sum(p["pay"] + p["scales"] for p in parts)
        return out, ev, sh

    def _gate(self, flag_ptr: int, seq: int, src_rank: int):
        """Host gate (lazy module loading only): return once flags[src] >= seq is visible."""
        ext = self._ext
        word = flag_ptr & ~7
        shift = 32 if (flag_ptr & 4) else 0
        deadline = time.monotonic() + self._wait_timeout_s
        pause = 5e-6
        while True:
            with self._gate_lock:
                v = (ext.read_u64s(word, 1, self._gate_stream)[0] >> shift) & 0xFFFFFFFF
            if ((v - (seq & 0xFFFFFFFF)) & 0xFFFFFFFF) < 0x80000000:
                return
            if time.monotonic() > deadline:
                logger.error(f"device receive from party #{src_rank} (message {seq}) timed out after "
                             f"{self._wait_timeout_s:.0f}s")
                self._status_host[0] = 1
                self.stats["forced_releases"] += 1
                ext.write_u32(flag_ptr, seq & 0xFFFFFFFF)
                return
            time.sleep(pause)
            pause = min(pause * 1.5, 5e-4)

    def _validate_meta(self, meta: dict, off: int, total: int):
        """The metadata frame comes from a peer: never build a view or launch a copy from numbers
        that do not describe a block inside that peer's region."""
        import torch

        if not (0 <= off and HEADER <= total and off % ALIGN == 0 and off + total <= self.region_bytes):
            raise ValueError(f"device message outside the arena region: off={off} total={total}")
        for p in meta["parts"]:
            toff, pay, pad, scales = int(p["toff"]), int(p["pay"]), int(p["pad"]), int(p["scales"])
            if not (HEADER <= toff and toff % 16 == 0 and 0 <= pay <= pad and 0 <= scales
                    and toff + pad + scales <= total):
                raise ValueError(f"device message part outside its block: {p}")
            dtype = getattr(torch, str(p["dtype"]), None)
            if not isinstance(dtype, torch.dtype):
                raise ValueError(f"unknown dtype in device message: {p['dtype']!r}")
            numel = 1
            for d in p["shape"]:
                numel *= int(d)
            ne, e0 = int(p["ne"]), int(p["e0"])
            if numel != int(p["numel"]) or not (0 <= e0 and 0 <= ne and e0 + ne <= numel):
                raise ValueError(f"inconsistent tensor geometry in device message: {p}")
            if p["whole"] and (e0 != 0 or ne != numel):
                raise ValueError("a whole-tensor part must cover the tensor")
            itemsize = dtype.itemsize
            mode = int(p["mode"])
            want = ne * itemsize if mode == K.MODE_COPY else (ne * 2 if mode == K.MODE_F32_TO_BF16 else ne)
            if pay != want or (mode == K.MODE_BF16_TO_FP8 and scales != ne // 32) or \
                    mode not in (K.MODE_COPY, K.MODE_F32_TO_BF16, K.MODE_BF16_TO_FP8):
                raise ValueError(f"payload size does not match the tensor geometry: {p}")

    # -- flow control + watchdog (off the data path) --------------------------------------
    def _credit_loop(self):
        import torch

        ext = self._ext
        try:
            torch.cuda.set_device(self.device_index)
        except Exception:  # noqa: BLE001
            return
        while not self._stopping:
            self._credit_wake.wait(0.5)
            self._credit_wake.clear()
            try:
                while self._credit_q:
                    src, bid, view_bytes, bcast, ev = self._credit_q.popleft()
                    cs = self._credit_stream
                    if ev is None:
                        # a zero-copy block died on the host: order the release after whatever the
                        # consumers already enqueued that reads it
                        streams = list(self._consumer_streams()) + [torch.cuda.current_stream(self._dev).cuda_stream,
                                                                   torch.cuda.default_stream(self._dev).cuda_stream]
                        for s in dict.fromkeys(streams):
                            e = ext.event_create(False)
                            ext.event_record(e, s)
                            ext.stream_wait_event(cs, e)
                            ext.event_destroy(e)  # released by the driver once it has fired
                        with self._recv_lock:
                            self._view_bytes[src] -= view_bytes
                    else:
                        ext.stream_wait_event(cs, ev.handle)
                    idx_list = self._bfreed_idx if bcast else self._freed_idx
                    idx = idx_list[src]
                    idx_list[src] = idx + 1
                    ring = (OFF_BFREED_RING if bcast else OFF_FREED_RING) + self.group.rank * FREED_RING * 8
                    cnt = (OFF_BFREED_CNT if bcast else OFF_FREED_CNT) + 8 * self.group.rank
                    ext.credit(self._peer(src, ring + (idx % FREED_RING) * 8), bid, self._peer(src, cnt), idx + 1, cs)
                    K.launch_counter.add("fed_credit")
                    self.stats["blocks_released"] += 1
                self._watchdog()
            except Exception:  # noqa: BLE001
                if not self._stopping:
                    logger.exception("device channel credit thread")

    def _mark_consumers(self, tensor):
        """A copy-out tensor is allocated on the receive stream but read on the consumers'
        streams: tell the caching allocator, so the memory is not reused under a running kernel."""
        import torch

        for h in self._consumer_streams():
            ts = self._tstream_cache.get(h)
            if ts is None:
                ts = self._tstream_cache[h] = torch.cuda.ExternalStream(h, device=self._dev)
            tensor.record_stream(ts)
        tensor.record_stream(torch.cuda.default_stream(self._dev))

    def _watchdog(self):
        """A wait-value cannot time out by itself.  If a producer never raises its flag (the party
        died, its producer failed after the metadata went out) the host completes the wait with
        the error status, so streams unblock and ``fed.get`` raises instead of hanging."""
        ext = self._ext
        now = time.monotonic()
        keep = []
        while self._pending_waits:
            item = self._pending_waits.popleft()
            ev, t0, src, seq, bcast = item
            if ext.event_query(ev.handle):
                continue
            if now - t0 > self._wait_timeout_s:
                logger.error(f"device receive from party #{src} (message {seq}) timed out after "
                             f"{self._wait_timeout_s:.0f}s: releasing the stream with an error status")
                self._status_host[0] = 1
                self.stats["forced_releases"] += 1
                flag = self._local((OFF_BFLAGS if bcast else OFF_FLAGS) + 4 * src)
                ext.write_u32(flag, seq & 0xFFFFFFFF)
                item = (ev, now, src, seq, bcast)
            keep.append(item)
        self._pending_waits.extend(keep)

    # -- misc --------------------------------------------------------------------------
    def read_descriptor(self, src_rank: int, meta: dict) -> dict:
        """Debug/tests: the descriptor the producer's kernel left in the arena (D2H read)."""
        region = (self._barena_off if meta.get("bcast") else OFF_ARENA) + src_rank * self.region_bytes
        return unpack_desc(self._ext.read_bytes(self._local(region + meta["off"]), 64))

    def device_barrier(self, stream=None):
        """K5 on the given (default: current) stream."""
        K.cuda_ext().set_device(self.device_index)
        K.barrier(n_parties=self.group.size, my_rank=self.group.rank, sig_ptr=self.signal_ptr,
                  sig_peers=self.signal_peers, epoch=self.next_epoch(1), status_ptr=self._status_host.data_ptr(),
                  timeout_ns=self._timeout_ns, stream=stream)

    def check_status(self) -> int:
        """0 = healthy; 1 = a device-side receive wait timed out; 2 = a party barrier did;
        3 = a message descriptor did not match what the host frame announced."""
        return int(self._status_host[0])

    def raise_if_failed(self):
        st = self.check_status()
        if st:
            from ...exceptions import FedError

            what = {1: "a device-side wait for a peer's push timed out — the peer party is gone or its producer "
                       "never launched",
                    2: "a party barrier timed out",
                    3: "a message descriptor in the arena does not match the announced one — the parties' "
                       "controllers have diverged or a foreign job wrote into this arena"}.get(st, "unknown")
            raise FedError(f"NVLink device channel failure (status {st}): {what} "
                           f"(timeout {self._timeout_ns / 1e9:.0f}s).")

    def synchronize(self):
        self._push_stream.synchronize()
        for h in self._recv_streams:
            if h:
                self._ext.stream_synchronize(h)
        self._ext.stream_synchronize(self._credit_stream)

    def close(self):
        self._stopping = True
        self._credit_wake.set()
        self._credit_thread.join(timeout=2)
