"""The NVLink proxies' host logic on CPU: framing (`fedcuda` / `fedcuda1` envelopes, `__devq__`
metadata frames), message planning, ordering, chunk re-assembly, broadcast routing.

The device channel is replaced by a fake with the same interface that keeps "arena blocks" in a
dict: `plan_messages`, `layout_parts`, `part_meta` and the receiver-side `_validate_meta` are the
real `DeviceChannel` code; only the two methods that launch kernels (`push_message`,
`receive_message`) are re-implemented over host memory.  What this covers is what the reference
does in `fed/proxy/grpc/grpc_proxy.py:193-220,272-342` (serialise, key, deliver, decode) for objects
that contain device tensors — without needing a GPU to run.
"""
import threading
import types
from collections import OrderedDict

import pytest
import torch

from party_utils import comm_config, free_ports
from rayfed_b200.proxy.nvlink import channel as C
from rayfed_b200.proxy.nvlink import nvlink_proxy as NP


class FakeCuda(torch.Tensor):
    """A host tensor that says it lives on a device (the pickler diverts `is_cuda` tensors)."""

    is_cuda = property(lambda self: True)

    def cpu(self, *args, **kwargs):        # the D2H copy of the small-tensor path yields a host tensor
        return self.as_subclass(torch.Tensor).clone()


def fake_cuda(t: torch.Tensor) -> torch.Tensor:
    return torch.Tensor._make_subclass(FakeCuda, t)


class _Event:
    handle = 0


class FakeChannel:
    """DeviceChannel stand-in: same planning and validation code, blocks kept in a dict."""

    device_index = 0
    wire_dtype = None

    def __init__(self, parties, block_bytes=1 << 20, multicast=True):
        self.group = types.SimpleNamespace(parties=list(parties), size=len(parties), rank=0)
        self.region_bytes = 4 * block_bytes
        self.max_message_bytes = block_bytes
        self.can_multicast = multicast
        self.blocks = {}          # (src_rank, dest_rank | "bcast", seq) -> {part index: bytes}
        self.seq = {}
        self.pushed = []          # (dest_rank | None, n_parts) per push_message call
        self.lock = threading.Lock()
        self.sender_rank = 0      # set by the test: which party is "this process" when pushing

    _plan_tensor = C.DeviceChannel._plan_tensor
    plan_messages = C.DeviceChannel.plan_messages
    can_bcast = C.DeviceChannel.can_bcast
    _validate_meta = C.DeviceChannel._validate_meta

    def push_message(self, dest_rank, parts, key, ready_event=0, producer_stream=0):
        total = C.DeviceChannel.layout_parts(parts)
        where = "bcast" if dest_rank is None else dest_rank
        with self.lock:
            seq = self.seq[(self.sender_rank, where)] = self.seq.get((self.sender_rank, where), 0) + 1
            payload = {}
            for i, p in enumerate(parts):
                t = p["tensor"].contiguous().view(-1)
                payload[i] = t[p["e0"]: p["e0"] + p["ne"]].clone()
            self.blocks[(self.sender_rank, where, seq)] = payload
            self.pushed.append((dest_rank, len(parts)))
        metas = [C.DeviceChannel.part_meta(p) for p in parts]
        return {"bid": seq, "off": 0, "seq": seq, "total": total, "bcast": dest_rank is None, "parts": metas,
                "desc": ""}

    def receive_message(self, src_rank, meta, assembling):
        self._validate_meta(meta, int(meta["off"]), int(meta["total"]))
        me = self.receiver_rank()
        key = (src_rank, "bcast" if meta.get("bcast") else me, int(meta["seq"]))
        with self.lock:
            payload = self.blocks[key] if meta.get("bcast") else self.blocks.pop(key)
        out = []
        for i, p in enumerate(meta["parts"]):
            data = torch.as_tensor(payload[i])
            dtype = getattr(torch, p["dtype"])
            if p["whole"]:
                out.append((p["t"], data.view(dtype).view(p["shape"])))
                continue
            full = assembling.get(p["t"])
            if full is None:
                full = assembling[p["t"]] = torch.empty(p["numel"], dtype=dtype)
            full[p["e0"]: p["e0"] + p["ne"]] = data.view(dtype)
            if p["e0"] + p["ne"] >= p["numel"]:
                assembling.pop(p["t"], None)
                out.append((p["t"], full.view(p["shape"])))
            else:
                out.append((p["t"], None))
        return out, _Event(), 0

    # which party is receiving is decided by the thread (the receiver proxy's name is in it) or
    # by the test for the inline path
    _receiver = threading.local()

    def receiver_rank(self):
        return getattr(self._receiver, "rank", self.default_receiver)

    default_receiver = 1

    def close(self):
        pass


@pytest.fixture()
def rig(monkeypatch):
    """alice (sender) -> bob, carol (receivers) over the loopback native transport, one process."""
    parties = ["alice", "bob", "carol"]
    ports = free_ports(3)
    addresses = {p: f"127.0.0.1:{port}" for p, port in zip(parties, ports)}
    cfg = comm_config("native", device_min_bytes=1024)
    fake = FakeChannel(parties, block_bytes=1 << 20)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *_a, **_k: None)
    monkeypatch.setattr(NP.NvlinkReceiverProxy, "_rehome", lambda self, t: t)
    NP.attach_channel(fake)
    receivers = {}
    for p in ("bob", "carol"):
        rx = NP.NvlinkReceiverProxy(addresses[p], p, "job", {}, cfg)
        rx.start()
        assert rx.is_ready()[0]
        receivers[p] = rx

    # every receive thread must know which party it stands for
    orig_loop = NP.NvlinkReceiverProxy._receive_loop

    def loop(self, channel, src):
        FakeChannel._receiver.rank = parties.index(self._party)
        return orig_loop(self, channel, src)

    monkeypatch.setattr(NP.NvlinkReceiverProxy, "_receive_loop", loop)
    for rx in receivers.values():
        rx.start_unpackers(fake)
    tx = NP.NvlinkSenderProxy(addresses, "alice", "job", {}, cfg)
    yield types.SimpleNamespace(tx=tx, rx=receivers, fake=fake, parties=parties)
    NP._ChannelHolder.channel = None
    tx.stop()
    for rx in receivers.values():
        rx.stop()


def _get(rig, party, up, down):
    FakeChannel._receiver.rank = rig.parties.index(party)      # the consuming thread decodes inline frames
    return rig.rx[party].get_data("alice", up, down).result(10)


def test_single_block_objects_travel_inline(rig):
    w = fake_cuda(torch.arange(4096, dtype=torch.float32))
    b = fake_cuda(torch.ones(1024, dtype=torch.bfloat16))
    obj = OrderedDict(step=3, params={"w": w, "b": b}, note="hello", tiny=[1, 2.5, None])
    assert rig.tx.send("bob", obj, "1#0", "2").result(10) is True
    got = _get(rig, "bob", "1#0", "2")
    assert list(got) == ["step", "params", "note", "tiny"] and got["step"] == 3 and got["tiny"] == [1, 2.5, None]
    assert torch.equal(got["params"]["w"], torch.arange(4096, dtype=torch.float32))
    assert torch.equal(got["params"]["b"], torch.ones(1024, dtype=torch.bfloat16))
    assert rig.fake.pushed == [(1, 2)]              # ONE block for both tensors, no metadata frames
    # plain objects and tensors below `device_min_bytes` never touch the channel
    assert rig.tx.send("bob", {"k": [1, 2, 3]}, "3#0", "4").result(10) is True
    assert _get(rig, "bob", "3#0", "4") == {"k": [1, 2, 3]}
    small = fake_cuda(torch.arange(8, dtype=torch.int64))
    assert rig.tx.send("bob", (small, "x"), "5#0", "6").result(10) is True
    t, s = _get(rig, "bob", "5#0", "6")
    assert torch.equal(t, torch.arange(8)) and s == "x" and len(rig.fake.pushed) == 1


def test_oversize_tensors_are_chunked_and_reassembled_in_order(rig):
    big = fake_cuda(torch.arange(700_001, dtype=torch.float32))          # 2.8 MB through 1 MiB blocks
    side = fake_cuda(torch.full((2048,), 7, dtype=torch.int32))
    assert rig.tx.send("bob", [big, side, "tail"], "10#0", "11").result(10) is True
    # a second multi-block object right behind it: the receive thread works in frame order
    other = fake_cuda(torch.arange(600_000, dtype=torch.float32) * 2)
    assert rig.tx.send("bob", {"o": other}, "12#0", "13").result(10) is True
    got_big, got_side, tail = _get(rig, "bob", "10#0", "11")
    assert tail == "tail" and torch.equal(got_side, torch.full((2048,), 7, dtype=torch.int32))
    assert torch.equal(got_big, torch.arange(700_001, dtype=torch.float32))
    assert torch.equal(_get(rig, "bob", "12#0", "13")["o"], torch.arange(600_000, dtype=torch.float32) * 2)
    assert len(rig.fake.pushed) >= 3 + 1 + 3        # chunks of `big`, `side`, chunks of `other`
    assert not rig.fake.blocks                       # every unicast block was consumed


def test_broadcast_pushes_once_and_reaches_every_receiver(rig):
    w = fake_cuda(torch.arange(8192, dtype=torch.float32))
    acks = rig.tx.broadcast(["bob", "carol"], {"w": w, "round": 9}, "20#0", "21")
    assert {d: a.result(10) for d, a in acks.items()} == {"bob": True, "carol": True}
    assert rig.fake.pushed == [(None, 1)]            # one multicast push, not one per receiver
    for p in ("bob", "carol"):
        got = _get(rig, p, "20#0", "21")
        assert got["round"] == 9 and torch.equal(got["w"], torch.arange(8192, dtype=torch.float32))
    # no multicast mapping: the broadcast degrades to one unicast push per receiver
    rig.fake.can_multicast = False
    rig.fake.pushed.clear()
    acks = rig.tx.broadcast(["bob", "carol"], {"w": w}, "22#0", "23")
    assert all(a.result(10) for a in acks.values())
    assert sorted(rig.fake.pushed) == [(1, 1), (2, 1)]
    for p in ("bob", "carol"):
        assert torch.equal(_get(rig, p, "22#0", "23")["w"], torch.arange(8192, dtype=torch.float32))
    # objects without device tensors are serialised once and written to both rings
    acks = rig.tx.broadcast(["bob", "carol"], None, "24#0", "25")
    assert all(a.result(10) for a in acks.values())
    assert _get(rig, "bob", "24#0", "25") is None and _get(rig, "carol", "24#0", "25") is None


def test_batch_claims_decode_device_frames(rig):
    a = fake_cuda(torch.arange(2048, dtype=torch.float32))
    rig.tx.send("bob", a, "30#0", "33").result(10)
    rig.tx.send("bob", "plain", "31#0", "33").result(10)
    FakeChannel._receiver.rank = 1
    refs = rig.rx["bob"].get_data_many([("alice", "30#0"), ("alice", "31#0")], "33")
    assert torch.equal(refs[0].result(10), torch.arange(2048, dtype=torch.float32)) and refs[1].result(10) == "plain"
