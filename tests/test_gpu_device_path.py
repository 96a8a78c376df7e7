"""The host-free tensor path through the fed API (K1 push → descriptor + release flag →
cuStreamWaitValue32 on the consumer's receive stream → zero-copy arena views), its flow control,
chunking of tensors larger than an arena block, multi-tensor objects, ``num_returns`` on device
parties, the fused ``fed.aggregate`` and the overlap of a local task with an in-flight push.

One process per party; on a 1-GPU box the parties share the GPU."""
import os

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)
if torch.cuda.device_count() < 2 and os.environ.get("RAYFED_B200_TEST_SINGLE_GPU_PARTIES", "1") == "0":
    pytest.skip("multi-party GPU tests on one GPU were disabled", allow_module_level=True)

from party_utils import make_addresses, run_parties  # noqa: E402
from test_gpu_fed_nvlink import _init  # noqa: E402


def _zero_copy(party, addresses):
    import gc

    import rayfed_b200 as fed
    from rayfed_b200.proxy.nvlink import channel as C
    from rayfed_b200.proxy.nvlink import nvlink_proxy

    fed, dev = _init(party, addresses)
    held = {}

    @fed.remote
    def make(n, seed):
        g = torch.Generator(device="cuda").manual_seed(seed)
        return torch.randn(n, device="cuda", generator=g)

    @fed.remote
    def inspect(t, seed):
        ch = nvlink_proxy.current_channel()
        lo = ch.buf.local_ptr + C.OFF_ARENA
        hi = ch.buf.local_ptr + ch.buf.nbytes
        g = torch.Generator(device="cuda").manual_seed(seed)
        exp = torch.randn(t.numel(), device="cuda", generator=g)
        held["t"] = t  # keep the view alive past the task
        return {"in_arena": lo <= t.data_ptr() < hi, "equal": bool(torch.equal(t, exp)),
                "zero_copy_tensors": ch.stats["zero_copy_tensors"]}

    @fed.remote
    def drop():
        ch = nvlink_proxy.current_channel()
        before = ch.stats["blocks_released"]
        held.clear()
        gc.collect()
        import time

        deadline = time.time() + 10
        while ch.stats["blocks_released"] == before and time.time() < deadline:
            time.sleep(0.01)
        return ch.stats["blocks_released"] - before

    r = fed.get(inspect.party("bob").remote(make.party("alice").remote(1 << 20, 7), 7))  # 4 MiB: a view
    assert r["in_arena"] and r["equal"] and r["zero_copy_tensors"] >= 1, r
    small = fed.get(inspect.party("bob").remote(make.party("alice").remote(2048, 8), 8))  # 8 KiB: copied out
    assert small["equal"] and not small["in_arena"], small
    assert fed.get(drop.party("bob").remote()) >= 1  # the block goes back when the last view dies
    ch = nvlink_proxy.current_channel()
    assert ch.check_status() == 0
    fed.shutdown()


def test_received_tensors_are_zero_copy_views_of_the_arena():
    addresses = make_addresses(["alice", "bob"])
    run_parties(_zero_copy, ["alice", "bob"], args=(addresses,), timeout=120)


def _host_gated(party, addresses):
    """The fallback for lazy CUDA module loading: receives and the aggregate rendezvous are gated on
    the host (no stream wait is ever parked); same results, same zero-copy views."""
    os.environ["RAYFED_B200_DEVICE_WAIT"] = "host"
    _zero_copy(party, addresses)


def test_host_gated_receive_mode():
    addresses = make_addresses(["alice", "bob"])
    run_parties(_host_gated, ["alice", "bob"], args=(addresses,), timeout=120)


def _host_gated_aggregate(party, addresses):
    os.environ["RAYFED_B200_DEVICE_WAIT"] = "host"
    _fused_aggregate(party, addresses)


def test_host_gated_aggregate_rendezvous():
    addresses = make_addresses(["alice", "bob"])
    run_parties(_host_gated_aggregate, ["alice", "bob"], args=(addresses,), timeout=180)


def _descriptor(party, addresses):
    import rayfed_b200 as fed
    from rayfed_b200.proxy.nvlink import channel as C
    from rayfed_b200.proxy.nvlink import nvlink_proxy

    fed, dev = _init(party, addresses)
    job = "Anonymous_job"

    @fed.remote
    def make():
        return torch.arange(3 * 5 * 7 * 1024, device="cuda", dtype=torch.float32).reshape(3, 5, 7, 1024)

    @fed.remote
    def consume(t):
        return float(t.double().sum())

    obj = make.party("alice").remote()
    out = consume.party("bob").remote(obj)
    exp = float(torch.arange(3 * 5 * 7 * 1024, dtype=torch.float64).sum())
    assert fed.get(out) == exp
    ch = nvlink_proxy.current_channel()
    if party == "bob":
        # the descriptor the producer's kernel left in my arena (block 1 of alice's region, offset 0)
        d = ch.read_descriptor(0, {"off": 0})
        assert d["magic"] == C.DESC_MAGIC and d["seq"] == 1 and d["status"] == 1, d
        assert d["job_hash"] == C.fnv1a32(job), d
        assert d["up_task"] == 1 and d["up_idx"] == 0 and d["down"] == 2, d  # make = task 1, consume = task 2
        assert d["ndim"] == 4 and d["shape"] == [3, 5, 7, 1024] and d["n_tensors"] == 1, d
        assert d["nbytes"] == C.HEADER + 3 * 5 * 7 * 1024 * 4, d
    assert ch.check_status() == 0
    fed.shutdown()


def test_push_kernel_writes_the_device_descriptor():
    addresses = make_addresses(["alice", "bob"])
    run_parties(_descriptor, ["alice", "bob"], args=(addresses,), timeout=120)


def _chunked(party, addresses, wire):
    import rayfed_b200 as fed
    from rayfed_b200.proxy.nvlink import nvlink_proxy

    comm = {"device_arena_bytes": 16 << 20}
    if wire:
        comm["device_wire_dtype"] = wire
    fed, dev = _init(party, addresses, **comm)

    @fed.remote
    def make(n):
        return {"big": torch.arange(n, device="cuda", dtype=torch.float32) * 0.5, "fit": torch.ones(65536, device="cuda"),
                "big16": (torch.arange(n // 2, device="cuda") % 251).to(torch.bfloat16)}

    @fed.remote
    def check(d, n):
        ok_big = torch.allclose(d["big"], torch.arange(n, device="cuda", dtype=torch.float32) * 0.5, rtol=1e-2)
        if not wire:
            ok_big = ok_big and torch.equal(d["big"], torch.arange(n, device="cuda", dtype=torch.float32) * 0.5)
        ok16 = torch.equal(d["big16"], (torch.arange(n // 2, device="cuda") % 251).to(torch.bfloat16))
        return bool(ok_big) and bool(ok16) and float(d["fit"].sum())

    n = 12 << 20  # 48 MiB of fp32 through a 16 MiB region (4 MiB blocks): 12+ chunks
    ok = check.party("bob").remote(make.party("alice").remote(n), n)
    assert fed.get(ok) == 65536.0
    ch = nvlink_proxy.current_channel()
    if party == "alice":
        assert ch.stats["messages_sent"] >= 12, ch.stats  # chunked through the arena, not the host
    assert ch.check_status() == 0
    fed.shutdown()


@pytest.mark.parametrize("wire", [None, "bf16"])
def test_tensor_larger_than_arena_is_chunked_through_it(wire):
    addresses = make_addresses(["alice", "bob"])
    run_parties(_chunked, ["alice", "bob"], args=(addresses, wire), timeout=150)


def _held_views(party, addresses):
    """A consumer that keeps received tensors must not stall the producer: blocks are released out
    of order (free list, not a ring)."""
    import rayfed_b200 as fed
    from rayfed_b200.proxy.nvlink import nvlink_proxy

    fed, dev = _init(party, addresses, device_arena_bytes=32 << 20)

    @fed.remote
    class Sink:
        def __init__(self):
            self.keep = []
            self.total = 0.0

        def take(self, t, keep):
            if keep:
                self.keep.append(t)
            self.total += float(t[0])
            return len(self.keep)

        def result(self):
            return self.total, [float(k[0]) for k in self.keep]

    @fed.remote
    def make(i):
        return torch.full((1 << 20,), float(i), device="cuda")  # 4 MiB each

    sink = Sink.party("bob").remote()
    n = 40  # 160 MiB through a 32 MiB region while 3 blocks stay pinned for good
    for i in range(n):
        sink.take.remote(make.party("alice").remote(i), i in (1, 2, 3))
    total, kept = fed.get(sink.result.remote())
    assert total == float(sum(range(n))) and kept == [1.0, 2.0, 3.0]
    assert nvlink_proxy.current_channel().check_status() == 0
    fed.shutdown()


def test_consumer_holding_views_does_not_stall_the_producer():
    addresses = make_addresses(["alice", "bob"])
    run_parties(_held_views, ["alice", "bob"], args=(addresses,), timeout=150)


def _state_dict(party, addresses):
    import rayfed_b200 as fed
    from rayfed_b200.ops import kernels as K
    from rayfed_b200.proxy.nvlink import nvlink_proxy

    fed, dev = _init(party, addresses)

    @fed.remote
    def make():
        g = torch.Generator(device="cuda").manual_seed(3)
        return {f"layer{i}.weight": torch.randn(64 + i, 128, device="cuda", generator=g) for i in range(60)}

    @fed.remote
    def digest(sd):
        return sorted((k, float(v.double().sum())) for k, v in sd.items())

    got = fed.get(digest.party("bob").remote(make.party("alice").remote()))
    g = torch.Generator(device="cuda").manual_seed(3)
    exp = sorted((f"layer{i}.weight", float(torch.randn(64 + i, 128, device="cuda", generator=g).double().sum()))
                 for i in range(60))
    # generation order == dict order, so regenerate in order then sort
    assert got == exp
    ch = nvlink_proxy.current_channel()
    if party == "alice":
        # 60 tensors = 2 blocks (<= 48 segments per launch), not 60 messages
        assert ch.stats["tensors_sent"] == 60 and ch.stats["messages_sent"] <= 3, ch.stats
        assert K.launch_counter.by_kernel.get("fed_push_multi", 0) >= 1
    fed.shutdown()


def test_state_dict_travels_as_one_multi_tensor_block():
    addresses = make_addresses(["alice", "bob"])
    run_parties(_state_dict, ["alice", "bob"], args=(addresses,), timeout=120)


def _num_returns(party, addresses):
    import rayfed_b200 as fed

    fed, dev = _init(party, addresses)

    @fed.remote
    def two(n):
        return torch.ones(n, device="cuda"), torch.full((n,), 2.0, device="cuda")

    @fed.remote
    class A:
        def pair(self, x):
            return x * 3, x * 4

    @fed.remote
    def add(a, b):
        return float((a + b).sum())

    x, y = two.party("alice").options(num_returns=2).remote(4096)
    assert fed.get(add.party("bob").remote(x, y)) == 3.0 * 4096
    a = A.party("bob").remote()
    p, q = a.pair.options(num_returns=2).remote(x)
    assert fed.get(add.party("alice").remote(p, q)) == 7.0 * 4096
    fed.shutdown()


def test_num_returns_on_device_parties():
    addresses = make_addresses(["alice", "bob"])
    run_parties(_num_returns, ["alice", "bob"], args=(addresses,), timeout=120)


def _overlap(party, addresses):
    """A local @fed.remote task on its own lane runs while a push towards this party is still in
    flight: the receive is a parked stream wait, not a blocked thread or an occupied SM."""
    import time

    import rayfed_b200 as fed

    fed, dev = _init(party, addresses, device_arena_bytes=256 << 20)

    @fed.remote
    class Slow:
        def big(self, n):
            x = torch.ones(n, device="cuda")
            torch.cuda._sleep(int(2.0e9))  # ~1 s of device time before the value is complete
            return x * 2

    @fed.remote
    class Local:
        def work(self):
            t0 = time.time()
            a = torch.ones(1 << 20, device="cuda")
            for _ in range(10):
                a = a * 1.0001
            torch.cuda.current_stream().synchronize()
            return time.time() - t0

    @fed.remote
    class Consumer:
        def use(self, t):
            return float(t[0])

    slow = Slow.party("alice").remote()
    local = Local.party("bob").remote()
    cons = Consumer.party("bob").remote()
    v = slow.big.remote(8 << 20)  # returns on the host at once; the device is busy for ~1 s
    got = cons.use.remote(v)  # bob: metadata arrives immediately, the receive stream parks on the flag
    time.sleep(0.2)
    t_local = fed.get(local.work.remote())  # issued while the push is pending
    t0 = time.time()
    assert fed.get(got) == 2.0
    if party == "bob":
        # had it waited for the push it would have taken the producer's full second of device time
        assert t_local < 0.8, f"the local task waited for the in-flight push ({t_local:.2f}s)"
    fed.shutdown()


def test_local_task_overlaps_an_in_flight_push():
    addresses = make_addresses(["alice", "bob"])
    run_parties(_overlap, ["alice", "bob"], args=(addresses,), timeout=120)


def _fused_aggregate(party, addresses):
    """fed.aggregate on the fused path: values never cross the DAG, the kernel averages them."""
    import rayfed_b200 as fed
    from rayfed_b200.ops import kernels as K

    fed, dev = _init(party, addresses)
    parties = list(addresses)
    n = len(parties)

    @fed.remote
    def make(rank):
        return {"w": torch.full((1000, 33), float(rank), device="cuda"),
                "b": torch.arange(64, device="cuda").bfloat16() * (rank + 1)}

    objs = [make.party(p).remote(i) for i, p in enumerate(parties)]
    before = fed._private.global_context.get_global_context().current_seq_id()
    new = fed.aggregate(objs)
    after = fed._private.global_context.get_global_context().current_seq_id()
    assert after - before == 2 * n + n  # n hidden actors (2 ids each) + n mean() calls
    assert [o.get_party() for o in new] == parties
    got = fed.get(new)  # only now do the (identical) means travel
    m = (n - 1) / 2.0
    for g in got:
        assert g["w"].dtype == torch.float32 and torch.allclose(g["w"].cpu(), torch.full((1000, 33), m))
        exp_b = (torch.arange(64).float() * ((n + 1) / 2.0)).bfloat16()
        assert g["b"].dtype == torch.bfloat16 and torch.allclose(g["b"].float().cpu(), exp_b.float(), rtol=2e-2)
    new2 = fed.aggregate([make.party(p).remote(i + 1) for i, p in enumerate(parties)])  # cached slot, second epoch
    assert torch.allclose(fed.get(new2[0])["w"].cpu(), torch.full((1000, 33), m + 1.0))
    assert K.launch_counter.by_kernel.get("fed_aggregate", 0) >= 2
    fed.shutdown()


def test_fed_aggregate_fused_path_two_parties():
    addresses = make_addresses(["alice", "bob"])
    run_parties(_fused_aggregate, ["alice", "bob"], args=(addresses,), timeout=180)


def _mixed_cpu_gpu(party, addresses):
    """a CPU-only peer must not hang the GPU party's init; both fall back to host frames."""
    import rayfed_b200 as fed
    from party_utils import comm_config
    from rayfed_b200.proxy.nvlink import nvlink_proxy

    cfg = {"cross_silo_comm": comm_config("auto", timeout_in_ms=20000)}
    if party == "alice":
        cfg["device"] = "cuda:0"
    fed.init(addresses=addresses, party=party, logging_level="warning", config=cfg)

    @fed.remote
    def make():
        return torch.arange(100000, device="cuda" if torch.cuda.is_available() else "cpu", dtype=torch.float32)

    @fed.remote
    def total(t):
        return float(t.double().sum())

    v = fed.get(total.party("bob").remote(make.party("alice").remote()))
    assert v == float(torch.arange(100000, dtype=torch.float64).sum())
    assert nvlink_proxy.current_channel() is None  # negotiated away
    fed.shutdown()


def test_mixed_cpu_and_gpu_parties_fall_back_to_the_host_path():
    addresses = make_addresses(["alice", "bob"])
    run_parties(_mixed_cpu_gpu, ["alice", "bob"], args=(addresses,), timeout=120)


def _agg_behind_pushes(party, addresses):
    """Regression (N=8 bench hang, round 2): the owner of a broadcast races ahead and launches the
    cross-party aggregate while its pushes to the last receivers are still queued.  The aggregate
    kernel spins in its barriers with every SM occupied; without the SM-free rendezvous in front of
    it those pushes starve, the last receivers never reach the aggregate, and everybody deadlocks."""
    import rayfed_b200 as fed
    from rayfed_b200.fl import FedAggregator, FlatBucket, device_context

    fed, dev = _init(party, addresses, device_arena_bytes=512 << 20)
    parties = list(addresses)

    @fed.remote
    class W:
        def __init__(self, rank):
            self.group, self.heap, self.channel = device_context()
            m = torch.nn.Linear(4096, 4096, bias=False, device="cuda", dtype=torch.bfloat16)
            with torch.no_grad():
                m.weight.fill_(float(rank + 1))
            self.m = m
            self.bucket = FlatBucket(m, self.heap, name="w", n_parties=self.group.size)
            self.agg = FedAggregator(self.bucket, self.group, self.channel, kind="fedavg", heap=self.heap)
            self.big = torch.ones(96 << 20, dtype=torch.uint8, device="cuda")

        def burst(self):
            return self.big  # 96 MiB, pushed to every other party by fed.get

        def aggregate(self):
            self.agg.aggregate()
            torch.cuda.current_stream().synchronize()
            return float(self.bucket.flat_w.float().mean())

    ws = {p: W.party(p).remote(i) for i, p in enumerate(parties)}
    for rep in range(3):
        for _ in range(3):
            got = fed.get(ws[parties[0]].burst.remote())  # owner returns at once; pushes still queued
            assert got.numel() == 96 << 20
        means = fed.get([ws[p].aggregate.remote() for p in parties])
        assert means == [1.5, 1.5] or all(abs(m - means[0]) < 1e-6 for m in means), means
    fed.shutdown()


def test_aggregate_launched_behind_queued_pushes_does_not_deadlock():
    addresses = make_addresses(["alice", "bob"])
    run_parties(_agg_behind_pushes, ["alice", "bob"], args=(addresses,), timeout=180)
