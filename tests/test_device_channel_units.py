"""CPU-side unit and property tests of the device channel's host logic: the free-list arena
allocator (sender side), the 64-byte descriptor codec and the message planner (multi-tensor
blocks, chunking).  The kernels and the stream path are covered by the GPU suites."""
import random
import types

import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

torch = pytest.importorskip("torch")

from rayfed_b200.ops import kernels as K  # noqa: E402
from rayfed_b200.proxy.nvlink import channel as C  # noqa: E402


def _check_invariants(alloc: C.RegionAllocator):
    spans = sorted([(o, n) for o, n in alloc._blocks.values()] + [(f[0], f[1]) for f in alloc._free])
    pos = 0
    for off, ln in spans:
        assert off == pos and ln > 0, (spans, pos)  # no gap, no overlap
        pos = off + ln
    assert pos == alloc.size
    free = sorted(alloc._free)
    for a, b in zip(free, free[1:]):
        assert a[0] + a[1] < b[0], "adjacent free intervals must be coalesced"
    for off, ln in alloc._blocks.values():
        assert off % C.ALIGN == 0 and ln % C.ALIGN == 0


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(1, 300_000), st.floats(0, 1)), min_size=1, max_size=200), st.integers(0, 2**31))
def test_region_allocator_never_overlaps_and_coalesces(ops, seed):
    rng = random.Random(seed)
    alloc = C.RegionAllocator(1 << 20)
    live = []
    for nbytes, p_free in ops:
        got = alloc.try_alloc(nbytes)
        if got is not None:
            bid, off = got
            assert off % C.ALIGN == 0 and off + nbytes <= alloc.size
            live.append(bid)
        while live and rng.random() < p_free:
            bid = live.pop(rng.randrange(len(live)))  # out-of-order release
            assert alloc.free(bid) and not alloc.free(bid)
        _check_invariants(alloc)
    for bid in live:
        alloc.free(bid)
    _check_invariants(alloc)
    assert alloc._free == [[0, alloc.size]] and alloc.outstanding == 0


def test_region_allocator_behaves_like_a_ring_and_skips_pinned_blocks():
    alloc = C.RegionAllocator(1 << 20)
    a = alloc.try_alloc(256 << 10)
    b = alloc.try_alloc(256 << 10)  # this one stays pinned (a consumer keeps a view)
    c = alloc.try_alloc(256 << 10)
    d = alloc.try_alloc(256 << 10)
    assert [x[1] for x in (a, b, c, d)] == [0, 256 << 10, 512 << 10, 768 << 10]
    assert alloc.try_alloc(1) is None  # full
    alloc.free(a[0])
    alloc.free(c[0])
    e = alloc.try_alloc(200 << 10)  # wraps to the start
    f = alloc.try_alloc(200 << 10)  # skips the pinned block b
    assert e[1] == 0 and f[1] == 512 << 10
    assert alloc.try_alloc(300 << 10) is None  # no hole is big enough, although 300 KiB are free in total
    alloc.free(b[0])
    alloc.free(e[0])
    g = alloc.try_alloc(500 << 10)
    assert g is not None and g[1] == 0


def test_descriptor_codec_roundtrip_and_key_words():
    raw = C.pack_desc(seq=7, job_hash=C.fnv1a32("job"), key=C._key_words("12#3", 45), offset=4096, nbytes=1 << 20,
                      dtype="bfloat16", mode=2, ndim=3, n_tensors=5, shape=(2, 3, 4))
    assert len(raw) == 64
    d = C.unpack_desc(raw)
    assert d["magic"] == C.DESC_MAGIC and d["seq"] == 7 and d["status"] == 1
    assert (d["up_task"], d["up_idx"], d["down"]) == (12, 3, 45)
    assert d["offset"] == 4096 and d["nbytes"] == 1 << 20 and d["mode"] == 2 and d["ndim"] == 3
    assert d["n_tensors"] == 5 and d["shape"] == [2, 3, 4, 0] and d["dtype"] == C._DTYPE_CODE["bfloat16"]
    # non-numeric ids (exchange tags, pings) are hashed, deterministically
    assert C._key_words("__xchg__/a", "ping") == C._key_words("__xchg__/a", "ping")
    assert C._key_words("__xchg__/a", "ping") != C._key_words("__xchg__/b", "ping")
    assert C.fnv1a32("") == 0x811C9DC5


def _planner(region_bytes, wire=None):
    stub = types.SimpleNamespace(region_bytes=region_bytes, wire_dtype=wire)
    stub.max_message_bytes = C.DeviceChannel.max_message_bytes.fget(stub)
    stub._plan_tensor = lambda t: C.DeviceChannel._plan_tensor(stub, t)
    return lambda tensors: C.DeviceChannel.plan_messages(stub, tensors)


def test_planner_packs_small_tensors_and_chunks_large_ones():
    plan = _planner(16 << 20)  # blocks of at most 4 MiB
    tensors = [torch.zeros(1000), torch.zeros(70_000), torch.zeros(3 << 20), torch.zeros(10)]  # 12 MiB in the middle
    msgs = plan(tensors)
    # the two small tensors share a block; the 12 MiB tensor is cut into >= 3 element ranges; the tail rides alone
    assert [p["t"] for p in msgs[0]] == [0, 1] and all(p["whole"] for p in msgs[0])
    chunks = [m[0] for m in msgs if not m[0]["whole"]]
    assert len(chunks) >= 3 and all(c["t"] == 2 for c in chunks)
    assert sum(c["ne"] for c in chunks) == 3 << 20
    assert [c["e0"] for c in chunks] == [sum(x["ne"] for x in chunks[:i]) for i in range(len(chunks))]
    assert all(c["e0"] % 256 == 0 for c in chunks)  # 16-byte aligned sources, whole fp8 scale blocks
    cap = (16 << 20) // 4 - C.HEADER
    for m in msgs:
        assert sum(C._pad(p["pay"]) + C._pad(p["scales"]) for p in m) <= cap
    assert [p["t"] for p in msgs[-1]] == [3]


def test_planner_respects_the_segment_limit_and_wire_formats():
    plan = _planner(256 << 20)
    msgs = plan([torch.zeros(2048) for _ in range(100)])
    assert [len(m) for m in msgs] == [48, 48, 4]  # <= 48 segments ride one launch
    plan16 = _planner(64 << 20, wire="bf16")
    (m,) = plan16([torch.zeros(4096, dtype=torch.float32)])
    assert m[0]["mode"] == K.MODE_F32_TO_BF16 and m[0]["pay"] == 4096 * 2
    plan8 = _planner(64 << 20, wire="fp8_e4m3")
    (m,) = plan8([torch.zeros(4096, dtype=torch.bfloat16)])
    assert m[0]["mode"] == K.MODE_BF16_TO_FP8 and m[0]["pay"] == 4096 and m[0]["scales"] == 4096 // 32
    (m,) = plan8([torch.zeros(100, dtype=torch.bfloat16)])  # not a multiple of 32: plain copy
    assert m[0]["mode"] == K.MODE_COPY


def test_receiver_validates_peer_supplied_metadata():
    stub = types.SimpleNamespace(region_bytes=1 << 20)
    check = lambda meta: C.DeviceChannel._validate_meta(stub, meta, int(meta["off"]), int(meta["total"]))  # noqa: E731
    good = {"off": 0, "total": C.HEADER + 4096, "parts": [
        {"t": 0, "e0": 0, "ne": 1024, "mode": K.MODE_COPY, "pay": 4096, "pad": 4096, "scales": 0, "toff": C.HEADER,
         "whole": True, "numel": 1024, "dtype": "float32", "shape": [32, 32]}]}
    check(good)
    import copy

    for mutate in (
        lambda m: m.update(off=(1 << 20) - 256),                 # block runs past the region
        lambda m: m.update(off=-256),
        lambda m: m["parts"][0].update(toff=64),                  # inside the header
        lambda m: m["parts"][0].update(toff=C.HEADER + 4096),     # past the block
        lambda m: m["parts"][0].update(shape=[64, 64]),           # geometry does not match numel
        lambda m: m["parts"][0].update(pay=8192, pad=8192),       # payload bigger than the tensor
        lambda m: m["parts"][0].update(dtype="os"),               # not a dtype
        lambda m: m["parts"][0].update(ne=512),                   # "whole" but partial
        lambda m: m["parts"][0].update(mode=9),
    ):
        bad = copy.deepcopy(good)
        mutate(bad)
        with pytest.raises(ValueError):
            check(bad)


@pytest.mark.parametrize("wire", [None, "bf16", "fp8_e4m3"])
def test_every_planned_message_passes_the_receivers_validation(wire):
    """Sender-side planning and receiver-side validation must agree for whole tensors, packed
    objects and chunked tensors in every wire format (CPU tensors stand in: only geometry matters)."""
    import torch

    region = 8 << 20
    stub = types.SimpleNamespace(wire_dtype=wire, max_message_bytes=max(region // 4, 1 << 20), region_bytes=region)
    stub._plan_tensor = lambda t: C.DeviceChannel._plan_tensor(stub, t)
    tensors = [torch.zeros(1024, dtype=torch.float32), torch.zeros(33, 7, dtype=torch.bfloat16),
               torch.zeros(4096, dtype=torch.bfloat16), torch.zeros(3 << 20, dtype=torch.bfloat16),
               torch.zeros(700_001, dtype=torch.float32), torch.zeros(5, dtype=torch.int64),
               torch.zeros(64, 64, dtype=torch.uint8), torch.zeros(16, dtype=torch.bool)]
    messages = C.DeviceChannel.plan_messages(stub, tensors)
    assert len(messages) > 2          # the two big tensors are chunked
    covered = {}
    for parts in messages:
        total = C.DeviceChannel.layout_parts(parts)
        assert total <= stub.max_message_bytes
        meta = {"off": 0, "total": total, "parts": [C.DeviceChannel.part_meta(p) for p in parts]}
        C.DeviceChannel._validate_meta(stub, meta, 0, total)
        for p in meta["parts"]:
            covered[p["t"]] = covered.get(p["t"], 0) + p["ne"]
    assert covered == {i: t.numel() for i, t in enumerate(tensors)}
