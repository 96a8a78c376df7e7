"""NVLink transport = native host frames for the object skeleton + the DeviceChannel
(K1 push / K4 stream wait / K1b multicast) for every CUDA tensor leaf.

What crosses the host socket for ``send(dest, obj)`` is a pickle of ``obj`` in which each large
CUDA tensor was replaced by a persistent id ``idx``, plus one small JSON *metadata* frame per
arena block (block id, offsets, shapes).  The tensor bytes never touch the host: the push kernel
stores them into the destination party's arena over NVLink.  Both frames leave at **launch**
time — the sender's host does not wait for the push kernel, and the receiver's host does not
wait for the data: a per-source thread orders that source's receive stream after the block's
release flag with ``cuStreamWaitValue32`` and hands the consumer zero-copy views of the arena
(or the output of an unpack kernel for wire formats / small / chunked tensors) together with
the CUDA event the consumer's stream has to wait for.

This is the B200-native replacement of the whole pickle→protobuf→HTTP/2 path for tensors
(SURVEY §3.4 — ≥5 host copies, 3 process hops); non-tensor objects keep the cheap host frame.

Security parity (fed/_private/serialization_utils.py:24-83): the skeleton is decoded by the
*same* restricted unpickler as any host payload when ``serializing_allowed_list`` is set; the
metadata frames are JSON, never pickle.
"""
from __future__ import annotations

import io
import json
import logging
import os as _os
import pickle
import threading
from typing import Dict

import cloudpickle

from ..._private.serialization import Deserializer, _RestrictedUnpickler
from ...config import CrossSiloMessageConfig
from ...core.runtime import DeviceValue, is_device_value
from ...exceptions import FedRemoteError, FedTaskError
from ..native.socket_proxy import NativeReceiverProxy, NativeSenderProxy

logger = logging.getLogger(__name__)

# K1b multicast push for owner -> everyone broadcasts (fed.get).  On by default where the
# channel got an NVLS mapping; RAYFED_B200_MC_PUSH=0 forces per-destination unicast pushes.
_MC_PUSH_ENABLED = _os.environ.get("RAYFED_B200_MC_PUSH", "1") != "0"
_DEVQ = "__devq__"
_TAG = "fedcuda"
_TAG1 = "fedcuda1"  # single-block object: metadata rides inside the data frame (no receive-thread hop)
_SMALL = "fedcuda-small"
# Objects that fit one arena block of at most this size are announced inside their own data frame and
# materialised by the consuming thread itself — two thread hand-overs less on the latency path.
# Bigger / multi-block objects keep the eager, in-order receive thread (chunk assembly, prompt
# copy-out and credit return).
_INLINE_MAX_BYTES = int(_os.environ.get("RAYFED_B200_INLINE_META_BYTES", str(4 << 20)))
_DEFAULT_MIN_BYTES = 4096


class _DevicePickler(cloudpickle.Pickler):
    """Pickles an object graph, diverting big CUDA tensors to ``sink``."""

    def __init__(self, file, sink, min_bytes, before_host_copy=None):
        super().__init__(file, protocol=5)
        self._sink = sink
        self._min_bytes = min_bytes
        self._before_host_copy = before_host_copy
        self.small_tensors = 0

    def persistent_id(self, obj):
        import torch

        if isinstance(obj, torch.Tensor) and obj.is_cuda and obj.layout == torch.strided:
            nbytes = obj.numel() * obj.element_size()
            if nbytes >= self._min_bytes:
                return (_TAG, self._sink(obj.detach()))
            # small tensors ride the host frame — as CPU data, re-homed on the *receiver's*
            # device (a plain pickle would resurrect them on the sender's device index and
            # make every party open a CUDA context on every other party's GPU)
            if self._before_host_copy is not None and self.small_tensors == 0:
                self._before_host_copy()  # the producer's stream must be done before a D2H read
            self.small_tensors += 1
            return (_SMALL, obj.detach().cpu())
        return None


def _persistent_load(self, pid):
    if isinstance(pid, tuple) and len(pid) == 2 and pid[0] == _TAG and isinstance(pid[1], int):
        return self._resolve(pid[1])
    if isinstance(pid, tuple) and len(pid) == 2 and pid[0] == _SMALL:
        return self._rehome(pid[1])
    raise pickle.UnpicklingError(f"unsupported persistent id {pid!r}")


class _DeviceUnpickler(pickle.Unpickler):
    """Skeleton unpickler without an allow-list."""

    persistent_load = _persistent_load


class _RestrictedDeviceUnpickler(_RestrictedUnpickler):
    """Skeleton unpickler when `serializing_allowed_list` is configured: same `find_class` rules as
    every other cross-party payload."""

    persistent_load = _persistent_load


def _make_unpickler(file, allowed, resolve, rehome):
    """Skeleton unpickler: the restricted one whenever an allow-list is configured.  (The two
    classes are module-level: building a class per message cost ~14 us on the receive path.)"""
    u = _DeviceUnpickler(file) if allowed is None else _RestrictedDeviceUnpickler(file, allowed)
    u._resolve = resolve
    u._rehome = rehome
    return u


class _ChannelHolder:
    """The DeviceChannel of this party, shared by its sender and receiver."""

    channel = None
    lock = threading.Lock()


def attach_channel(channel) -> None:
    _ChannelHolder.channel = channel


def detach_channel() -> None:
    ch, _ChannelHolder.channel = _ChannelHolder.channel, None
    if ch is not None:
        ch.close()


def current_channel():
    return _ChannelHolder.channel


def _split(data):
    """(value, event handle, producer stream handle, event owner)."""
    if is_device_value(data):
        return data.value, data.event, data.stream, data.owner
    return data, 0, 0, None


class NvlinkSenderProxy(NativeSenderProxy):
    DEVICE_AWARE = True

    def __init__(self, addresses: Dict, party: str, job_name: str, tls_config: Dict, proxy_config: Dict = None):
        super().__init__(addresses, party, job_name, tls_config, proxy_config)
        cfg: CrossSiloMessageConfig = self._proxy_config
        self._min_bytes = cfg.device_min_bytes if cfg.device_min_bytes is not None else _DEFAULT_MIN_BYTES
        self._order_lock = threading.Lock()
        self._k = {}  # dest -> number of metadata frames sent (= the receiver's processing order)

    def _serialise(self, data, owner):
        tensors = []

        def sink(t):
            tensors.append(t)
            return len(tensors) - 1

        buf = io.BytesIO()
        pickler = _DevicePickler(buf, sink, self._min_bytes, owner.device_sync if owner is not None else None)
        pickler.dump(data)
        return buf.getvalue(), tensors, pickler.small_tensors

    def _next_k(self, dest):
        k = self._k.get(dest, 0) + 1
        self._k[dest] = k
        return k

    def _meta_frame(self, dest, meta, last, n_tensors):
        k = self._next_k(dest)
        meta = dict(meta, last=bool(last), n=n_tensors)
        self._transport.send(dest, f"{_DEVQ}/{self._party}", str(k), json.dumps(meta).encode(), False)
        return k

    def send(self, dest_party, data, upstream_seq_id, downstream_seq_id):
        data, event, pstream, owner = _split(data)
        channel = current_channel()
        if channel is None or isinstance(data, FedRemoteError):
            if owner is not None:
                owner.device_sync()
            return super().send(dest_party, data, upstream_seq_id, downstream_seq_id)

        skeleton, tensors, small = self._serialise(data, owner)
        up, down = str(upstream_seq_id), str(downstream_seq_id)
        if not tensors:
            frame = pickle.dumps((_TAG, 0, skeleton), protocol=5) if small else skeleton
            return self._transport.send(dest_party, up, down, frame, False)
        dest_rank = channel.group.parties.index(dest_party)
        with self._order_lock:
            # one metadata frame per arena block, sent the moment its push kernel is *launched*
            messages = channel.plan_messages(tensors)
            if len(messages) == 1 and sum(p["pay"] + p["scales"] for p in messages[0]) <= _INLINE_MAX_BYTES:
                meta = channel.push_message(dest_rank, messages[0], (up, down), event, pstream)
                meta.update(last=True, n=len(tensors))
                frame = pickle.dumps((_TAG1, json.dumps(meta).encode(), skeleton), protocol=5)
                return self._transport.send(dest_party, up, down, frame, False)
            k = 0
            for i, parts in enumerate(messages):
                meta = channel.push_message(dest_rank, parts, (up, down), event, pstream)
                k = self._meta_frame(dest_party, meta, i == len(messages) - 1, len(tensors))
        frame = pickle.dumps((_TAG, k, skeleton), protocol=5)
        return self._transport.send(dest_party, up, down, frame, False)

    # -- K1b: owner -> all-parties broadcast (fed.get, weight redistribution) -------------------
    def can_broadcast(self) -> bool:
        return True  # serialise once in any case; multicast the tensors when the channel can

    def broadcast(self, dest_parties, data, upstream_seq_id, downstream_seq_id):
        """Send the same object to every party in ``dest_parties`` (all the others): each big
        CUDA tensor is pushed ONCE with the multicast kernel; the skeleton + metadata go
        to each destination.  Returns {dest: ack Ref}."""
        orig = data
        data, event, pstream, owner = _split(data)
        channel = current_channel()
        if channel is None or isinstance(data, FedRemoteError):
            if owner is not None:
                owner.device_sync()
            return super().broadcast(dest_parties, data, upstream_seq_id, downstream_seq_id)
        skeleton, tensors, small = self._serialise(data, owner)
        up, down = str(upstream_seq_id), str(downstream_seq_id)
        send = self._transport.send
        dest_parties = list(dest_parties)

        def send_all(frame):  # one GIL release around the N-1 writes
            return dict(zip(dest_parties, self._transport.send_many(dest_parties, up, down, frame, False)))

        if not tensors:
            return send_all(pickle.dumps((_TAG, 0, skeleton), protocol=5) if small else skeleton)
        if not (_MC_PUSH_ENABLED and channel.can_bcast(tensors)):
            return {d: self.send(d, orig, upstream_seq_id, downstream_seq_id) for d in dest_parties}
        ks = {}
        with self._order_lock:
            messages = channel.plan_messages(tensors)
            if len(messages) == 1 and sum(p["pay"] + p["scales"] for p in messages[0]) <= _INLINE_MAX_BYTES:
                meta = channel.push_message(None, messages[0], (up, down), event, pstream)
                meta.update(last=True, n=len(tensors))
                return send_all(pickle.dumps((_TAG1, json.dumps(meta).encode(), skeleton), protocol=5))
            for i, parts in enumerate(messages):
                meta = channel.push_message(None, parts, (up, down), event, pstream)
                for d in dest_parties:
                    ks[d] = self._meta_frame(d, meta, i == len(messages) - 1, len(tensors))
        return {d: send(d, up, down, pickle.dumps((_TAG, ks[d], skeleton), protocol=5), False)
                for d in dest_parties}


class NvlinkReceiverProxy(NativeReceiverProxy):
    def __init__(self, listening_address: str, party: str, job_name: str, tls_config: Dict,
                 proxy_config: Dict = None):
        super().__init__(listening_address, party, job_name, tls_config, proxy_config)
        self._loads = Deserializer(self._proxy_config.serializing_allowed_list)
        self._decoders = {}  # src party -> decoder closure
        self._ready = {}  # (src, k of the object's last block) -> (tensors, event, stream)
        self._cv = threading.Condition()
        self._threads = []
        self._waiting = {}  # src -> k the receive thread is parked on
        self._stopping = False
        self._timeout_s = (self._proxy_config.timeout_in_ms or 60000) / 1000.0

    # called by fed.init once the channel exists
    def start_unpackers(self, channel):
        for src in channel.group.parties:
            if src == self._party:
                continue
            t = threading.Thread(target=self._receive_loop, args=(channel, src), name=f"fed-recv-{src}",
                                 daemon=True)
            t.start()
            self._threads.append(t)

    def _receive_loop(self, channel, src: str):
        """Metadata frames of one source, in order.  Per frame: a stream wait-value, a
        descriptor check and (maybe) unpack kernels are *enqueued*; nothing here waits for the GPU."""
        import torch

        torch.cuda.set_device(channel.device_index)
        src_rank = channel.group.parties.index(src)
        k = 0
        tensors, assembling = {}, {}
        while not self._stopping:
            k += 1
            self._waiting[src] = k
            ref = self._mailbox.claim(f"{_DEVQ}/{src}", str(k))
            ref.wait(-1.0)
            if self._stopping:
                return
            try:
                raw = ref.result()
                if raw is None:
                    return
                meta = json.loads(raw)
                parts, ev, stream = channel.receive_message(src_rank, meta, assembling)
                for ti, t in parts:
                    if t is not None:
                        tensors[ti] = t
                if not meta["last"]:
                    continue
                out = ([tensors[i] for i in range(meta["n"])], ev, stream)
                tensors, assembling = {}, {}
            except BaseException as e:  # noqa: BLE001
                logger.exception("device receive failed")
                out = e
                tensors, assembling = {}, {}
            with self._cv:
                self._ready[(src, k)] = out
                self._cv.notify_all()

    def _rehome(self, cpu_tensor):
        ch = current_channel()
        if ch is None:
            return cpu_tensor
        return cpu_tensor.to(f"cuda:{ch.device_index}", non_blocking=False)

    def _take(self, src: str, k: int):
        with self._cv:
            ok = self._cv.wait_for(lambda: (src, k) in self._ready or self._stopping, self._timeout_s)
            if not ok or self._stopping:
                raise TimeoutError(f"device payload {k} from {src} never arrived")
            out = self._ready.pop((src, k))
        if isinstance(out, BaseException):
            raise out
        return out

    def _decoder_for(self, src_party):
        """Bytes -> value for messages of ``src_party`` (built once per source)."""
        dec = self._decoders.get(src_party)
        if dec is not None:
            return dec
        loads = self._loads

        def decode(data: bytes, is_error: bool):
            obj = loads(data)
            if isinstance(obj, tuple) and len(obj) == 3 and obj[0] == _TAG1:
                # single-block object: this (consuming) thread enqueues the stream wait itself
                channel = current_channel()
                if channel is None:
                    raise RuntimeError("received a device-channel message but this party has no device channel")
                meta = json.loads(obj[1])
                got, ev, stream = channel.receive_message(channel.group.parties.index(src_party), meta, {})
                tensors = dict(got)
                value = _make_unpickler(io.BytesIO(obj[2]), loads.allowed, lambda idx: tensors[idx],
                                        self._rehome).load()
                return DeviceValue(value, ev.handle, stream, ev)
            if isinstance(obj, tuple) and len(obj) == 3 and obj[0] == _TAG:
                _, k, skeleton = obj
                if k == 0:  # no device parts: only (possibly) small tensors to re-home
                    value = _make_unpickler(io.BytesIO(skeleton), loads.allowed, None, self._rehome).load()
                    if isinstance(value, FedRemoteError):
                        raise FedTaskError(value, "", "recv")
                    return value
                parts, ev, stream = self._take(src_party, k)
                value = _make_unpickler(io.BytesIO(skeleton), loads.allowed, lambda idx: parts[idx],
                                        self._rehome).load()
                # one event behind the source's receive stream orders every tensor of the object;
                # the consumer's stream waits for it (C++ executor) — the host does not
                return DeviceValue(value, ev.handle, stream, ev)
            if isinstance(obj, FedRemoteError):
                raise FedTaskError(obj, "", "recv")
            return obj

        self._decoders[src_party] = decode
        return decode

    def get_data(self, src_party, upstream_seq_id, curr_seq_id):
        ref = self._mailbox.claim(str(upstream_seq_id), str(curr_seq_id))
        ref.set_decoder(self._decoder_for(src_party))
        return ref

    def get_data_many(self, items, curr_seq_id):
        refs = self._mailbox.claim_many([str(up) for _, up in items], str(curr_seq_id))
        for (src, _), ref in zip(items, refs):
            ref.set_decoder(self._decoder_for(src))
        return refs

    def stop(self):
        self._stopping = True
        # unblock receive threads parked on their next metadata frame
        for src, k in list(self._waiting.items()):
            try:
                self._mailbox.deliver_obj(f"{_DEVQ}/{src}", str(k), None, False)
            except Exception:  # noqa: BLE001
                pass
        with self._cv:
            self._cv.notify_all()
        for t in self._threads:
            t.join(timeout=2)
        self._threads.clear()
        super().stop()
