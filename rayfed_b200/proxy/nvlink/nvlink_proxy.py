"""NVLink transport = native host frames for the object skeleton + the
DeviceChannel (K1 push / K4 unpack) for every CUDA tensor leaf.

What crosses the host socket for ``send(dest, obj)`` is a pickle of ``obj`` in
which each large CUDA tensor was replaced by a persistent id ``(k, idx)``; the
tensor bytes themselves never touch the host: they are stored by the push
kernel into the destination party's arena over NVLink.  Receive-side unpack
kernels are launched by a per-source ordered thread as soon as the descriptor
frame lands (arena regions are byte rings, so they must be consumed in FIFO
order), i.e. usually *before* the consumer task asks for the value; the
consumer just orders its CUDA stream after the unpack event.

This is the B200-native replacement of the whole pickle→protobuf→HTTP/2 path
for tensors (SURVEY §3.4 — ≥5 host copies, 3 process hops), while non-tensor
objects keep the cheap host frame.
"""
from __future__ import annotations

import io
import logging
import pickle
import queue
import threading
from typing import Dict

import cloudpickle

from ..._private.serialization import Deserializer
from ...config import CrossSiloMessageConfig
from ...core.runtime import _Produced
from ...exceptions import FedRemoteError, FedTaskError
from ..native.socket_proxy import NativeReceiverProxy, NativeSenderProxy, _ack_decoder

logger = logging.getLogger(__name__)

import os as _os

# K1b multicast push for owner -> everyone broadcasts (fed.get).  On by default where the
# channel got an NVLS mapping; RAYFED_B200_MC_PUSH=0 forces per-destination unicast pushes.
_MC_PUSH_ENABLED = _os.environ.get("RAYFED_B200_MC_PUSH", "1") != "0"
_DEVQ = "__devq__"
_TAG = "fedcuda"
_SMALL = "fedcuda-small"
_DEFAULT_MIN_BYTES = 4096


class _DevicePickler(cloudpickle.Pickler):
    """Pickles an object graph, diverting big CUDA tensors to ``sink``."""

    def __init__(self, file, sink, min_bytes, before_host_copy=None, max_bytes=None):
        super().__init__(file, protocol=5)
        self._sink = sink
        self._min_bytes = min_bytes
        self._max_bytes = max_bytes  # tensors that do not fit an arena region spill to the host frame
        self._before_host_copy = before_host_copy
        self.small_tensors = 0

    def persistent_id(self, obj):
        import torch

        if isinstance(obj, torch.Tensor) and obj.is_cuda and obj.layout == torch.strided:
            nbytes = obj.numel() * obj.element_size()
            if nbytes >= self._min_bytes and (self._max_bytes is None or nbytes + 4096 <= self._max_bytes):
                return (_TAG, self._sink(obj.detach()))
            # small (and oversize) tensors ride the host frame — as CPU data, re-homed on the *receiver's*
            # device (a plain pickle would resurrect them on the sender's device index and
            # make every party open a CUDA context on every other party's GPU)
            if self._before_host_copy is not None and self.small_tensors == 0:
                self._before_host_copy()  # the producer's stream must be done before a D2H read
            self.small_tensors += 1
            return (_SMALL, obj.detach().cpu())
        return None


class _DeviceUnpickler(pickle.Unpickler):
    def __init__(self, file, resolve, rehome=None):
        super().__init__(file)
        self._resolve = resolve
        self._rehome = rehome or (lambda t: t)

    def persistent_load(self, pid):
        if isinstance(pid, tuple) and pid and pid[0] == _TAG:
            return self._resolve(pid[1])
        if isinstance(pid, tuple) and pid and pid[0] == _SMALL:
            return self._rehome(pid[1])
        raise pickle.UnpicklingError(f"unsupported persistent id {pid!r}")


class _ChannelHolder:
    """The DeviceChannel of this party, shared by its sender and receiver."""

    channel = None
    lock = threading.Lock()


def attach_channel(channel) -> None:
    _ChannelHolder.channel = channel


def detach_channel() -> None:
    _ChannelHolder.channel = None


def current_channel():
    return _ChannelHolder.channel


class NvlinkSenderProxy(NativeSenderProxy):
    DEVICE_AWARE = True

    def __init__(self, addresses: Dict, party: str, job_name: str, tls_config: Dict, proxy_config: Dict = None):
        super().__init__(addresses, party, job_name, tls_config, proxy_config)
        cfg: CrossSiloMessageConfig = self._proxy_config
        self._min_bytes = cfg.device_min_bytes if cfg.device_min_bytes is not None else _DEFAULT_MIN_BYTES
        self._order_lock = threading.Lock()
        self._k = {}  # dest -> number of device messages sent
        # Descriptor frames are released only once the push kernel has *finished*:
        # the receiver then launches its unpack kernel against data that is already
        # there.  (An eagerly launched unpack kernel that spins on the flag can
        # deadlock: its GPU front-end queue may be shared with a later-submitted
        # kernel the producer transitively waits for — measured on B200, see
        # profiles/README.md.)
        self._done_q = queue.SimpleQueue()
        self._done_thread = threading.Thread(target=self._completion_loop, name="fed-push-done", daemon=True)
        self._done_thread.start()

    def _completion_loop(self):
        while True:
            item = self._done_q.get()
            if item is None:
                return
            ev, dest_party, k, blob = item
            try:
                ev.synchronize()  # releases the GIL
                self._transport.send(dest_party, f"{_DEVQ}/{self._party}", str(k), blob, False, _ack_decoder)
            except Exception:  # noqa: BLE001
                logger.exception("failed to publish device descriptors")

    def stop(self):
        self._done_q.put(None)
        self._done_thread.join(timeout=5)
        super().stop()

    def send(self, dest_party, data, upstream_seq_id, downstream_seq_id):
        ready_event = None
        producer_stream = None
        if type(data) is _Produced:
            ready_event, producer_stream, data = data.event, data.stream, data.value
        channel = current_channel()
        if channel is None or isinstance(data, FedRemoteError):
            if ready_event is not None:
                ready_event.synchronize()
            return super().send(dest_party, data, upstream_seq_id, downstream_seq_id)

        tensors = []

        def sink(t):
            tensors.append(t)
            return len(tensors) - 1

        buf = io.BytesIO()
        pickler = _DevicePickler(buf, sink, self._min_bytes,
                                 ready_event.synchronize if ready_event is not None else None,
                                 max_bytes=channel.region_bytes)
        pickler.dump(data)
        if not tensors:
            if pickler.small_tensors:
                frame = pickle.dumps((_TAG, 0, buf.getvalue()), protocol=5)
            else:
                frame = buf.getvalue()  # plain host object: nothing to re-home
            return self._transport.send(dest_party, str(upstream_seq_id), str(downstream_seq_id), frame,
                                        False, _ack_decoder)
        dest_rank = channel.group.parties.index(dest_party)
        with self._order_lock:
            k = self._k.get(dest_party, 0) + 1
            self._k[dest_party] = k
            descs = [channel.push_tensor(dest_rank, t, ready_event, producer_stream=producer_stream) for t in tensors]
            self._done_q.put((channel.record_push_event(), dest_party, k, pickle.dumps(descs)))
        frame = pickle.dumps((_TAG, k, buf.getvalue()), protocol=5)
        return self._transport.send(dest_party, str(upstream_seq_id), str(downstream_seq_id), frame, False,
                                    _ack_decoder)


    # -- K1b: owner -> all-parties broadcast (fed.get, weight redistribution) -------------------
    def can_broadcast(self) -> bool:
        return True  # serialise once in any case; multicast the tensors when the channel can

    def broadcast(self, dest_parties, data, upstream_seq_id, downstream_seq_id):
        """Send the same object to every party in ``dest_parties`` (all the others): each big
        CUDA tensor is pushed ONCE with the multicast kernel; the skeleton + descriptors go
        to each destination.  Returns {dest: ack Ref}."""
        ready_event = None
        if type(data) is _Produced:
            ready_event, data = data.event, data.value
        channel = current_channel()
        if channel is None or isinstance(data, FedRemoteError):
            if ready_event is not None:
                ready_event.synchronize()
            return super().broadcast(dest_parties, data, upstream_seq_id, downstream_seq_id)
        tensors = []

        def sink(t):
            tensors.append(t)
            return len(tensors) - 1

        buf = io.BytesIO()
        pickler = _DevicePickler(buf, sink, self._min_bytes,
                                 ready_event.synchronize if ready_event is not None else None,
                                 max_bytes=channel.region_bytes)
        pickler.dump(data)
        acks = {}
        if not tensors:
            frame = pickle.dumps((_TAG, 0, buf.getvalue()), protocol=5) if pickler.small_tensors else buf.getvalue()
            for d in dest_parties:
                acks[d] = self._transport.send(d, str(upstream_seq_id), str(downstream_seq_id), frame, False,
                                               _ack_decoder)
            return acks
        with self._order_lock:
            descs = None
            if _MC_PUSH_ENABLED and all(channel.can_bcast_tensor(t) for t in tensors):  # all or nothing: no ring leaks
                descs = [channel.push_tensor_bcast(t, ready_event) for t in tensors]
                if any(d is None for d in descs):
                    raise RuntimeError("multicast push refused a tensor it had accepted")
            if descs is not None:
                ev = channel.record_push_event()
                blob = pickle.dumps(descs)
                ks = {}
                for d in dest_parties:
                    ks[d] = self._k.get(d, 0) + 1
                    self._k[d] = ks[d]
                    self._done_q.put((ev, d, ks[d], blob))
        if descs is None:
            data_again = _Produced(data, ready_event) if ready_event is not None else data
            return {d: self.send(d, data_again, upstream_seq_id, downstream_seq_id) for d in dest_parties}
        for d in dest_parties:
            frame = pickle.dumps((_TAG, ks[d], buf.getvalue()), protocol=5)
            acks[d] = self._transport.send(d, str(upstream_seq_id), str(downstream_seq_id), frame, False,
                                           _ack_decoder)
        return acks


class NvlinkReceiverProxy(NativeReceiverProxy):
    def __init__(self, listening_address: str, party: str, job_name: str, tls_config: Dict,
                 proxy_config: Dict = None):
        super().__init__(listening_address, party, job_name, tls_config, proxy_config)
        self._loads = Deserializer(self._proxy_config.serializing_allowed_list)
        self._unpacked = {}  # (src, k) -> list[(tensor, event)]
        self._cv = threading.Condition()
        self._threads = []
        self._waiting = {}  # src -> k the unpacker is parked on
        self._stopping = False

    # called by fed.init once the channel exists
    def start_unpackers(self, channel):
        for src in channel.group.parties:
            if src == self._party:
                continue
            t = threading.Thread(target=self._unpack_loop, args=(channel, src), name=f"fed-unpack-{src}",
                                 daemon=True)
            t.start()
            self._threads.append(t)

    def _unpack_loop(self, channel, src: str):
        import torch

        torch.cuda.set_device(channel.device_index)
        src_rank = channel.group.parties.index(src)
        k = 0
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
                descs = pickle.loads(raw)
                out = [channel.unpack_tensor(src_rank, d) for d in descs]
            except BaseException as e:  # noqa: BLE001
                logger.exception("device unpack failed")
                out = e
            with self._cv:
                self._unpacked[(src, k)] = out
                self._cv.notify_all()

    def _rehome(self, cpu_tensor):
        ch = current_channel()
        if ch is None:
            return cpu_tensor
        return cpu_tensor.to(f"cuda:{ch.device_index}", non_blocking=False)

    def _take(self, src: str, k: int, timeout: float = 120.0):
        with self._cv:
            ok = self._cv.wait_for(lambda: (src, k) in self._unpacked or self._stopping, timeout)
            if not ok or self._stopping:
                raise TimeoutError(f"device payload {k} from {src} never arrived")
            out = self._unpacked.pop((src, k))
        if isinstance(out, BaseException):
            raise out
        return out

    def get_data(self, src_party, upstream_seq_id, curr_seq_id):
        ref = self._mailbox.claim(str(upstream_seq_id), str(curr_seq_id))
        loads = self._loads

        def decode(data: bytes, is_error: bool):
            obj = loads(data)
            if isinstance(obj, tuple) and len(obj) == 3 and obj[0] == _TAG:
                _, k, skeleton = obj
                if k == 0:  # no device parts: only (possibly) small tensors to re-home
                    value = _DeviceUnpickler(io.BytesIO(skeleton), None, self._rehome).load()
                    if isinstance(value, FedRemoteError):
                        raise FedTaskError(value, "", "recv")
                    return value
                parts = self._take(src_party, k)
                value = _DeviceUnpickler(io.BytesIO(skeleton), lambda idx: parts[idx][0], self._rehome).load()
                # the receive stream is in-order: the last unpack's event covers them all
                return _Produced(value, parts[-1][1])
            if isinstance(obj, FedRemoteError):
                raise FedTaskError(obj, "", "recv")
            return obj

        ref.set_decoder(decode)
        return ref

    def stop(self):
        self._stopping = True
        # unblock unpackers parked on their next descriptor frame
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
