"""Multi-GPU tests written after this round's GPU budget was spent: they have NOT been run on
hardware yet.  The module name sorts last on purpose, so that with ``pytest -x`` a surprise
here cannot hide the results of the validated suites."""
import os

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)
if os.environ.get("RAYFED_B200_NEXT_ROUND_TESTS") != "1":
    pytest.skip("not yet executed on hardware: opt in with RAYFED_B200_NEXT_ROUND_TESTS=1 "
                "(tools/round_first_call.sh and tools/round_multi_gpu_call.sh do)", allow_module_level=True)
if torch.cuda.device_count() < 2 and os.environ.get("RAYFED_B200_TEST_SINGLE_GPU_PARTIES") != "1":
    pytest.skip("multi-party GPU tests need >= 2 GPUs", allow_module_level=True)

from party_utils import make_addresses, run_parties  # noqa: E402
from test_gpu_fed_nvlink import _init  # noqa: E402


def _oversize(party, addresses):
    """A tensor larger than an arena region must not fail the job: it spills to the host frame
     while tensors that fit keep using the device path."""
    import rayfed_b200 as fed
    from rayfed_b200.proxy.nvlink import nvlink_proxy

    fed, dev = _init(party, addresses, device_arena_bytes=1 << 20)

    @fed.remote
    def make(n):
        return {"big": torch.arange(n, device="cuda", dtype=torch.float32), "fit": torch.ones(65536, device="cuda")}

    @fed.remote
    def check(d, n):
        assert d["big"].is_cuda and d["fit"].is_cuda
        return bool(torch.equal(d["big"], torch.arange(n, device="cuda", dtype=torch.float32))) and float(d["fit"].sum())

    n = 1 << 20  # 4 MiB > the 1 MiB region
    ok = check.party("bob").remote(make.party("alice").remote(n), n)
    assert fed.get(ok) == 65536.0
    ch = nvlink_proxy.current_channel()
    if party == "alice":
        assert ch.stats["tensors_sent"] == 1, ch.stats  # only the tensor that fits went through the arena
    fed.shutdown()


def test_tensor_larger_than_arena_region_spills_to_host_path():
    addresses = make_addresses(["alice", "bob"])
    run_parties(_oversize, ["alice", "bob"], args=(addresses,), timeout=120)


def _fused_aggregate(party, addresses):
    """fed.aggregate on the fused path: values never cross the DAG, the kernel averages them."""
    import rayfed_b200 as fed
    from rayfed_b200.ops import kernels as K

    fed, dev = _init(party, addresses)
    parties = list(addresses)
    n = len(parties)

    @fed.remote
    def make(rank):
        return {"w": torch.full((1000, 33), float(rank), device="cuda"), "b": torch.arange(64, device="cuda").bfloat16() * (rank + 1)}

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
