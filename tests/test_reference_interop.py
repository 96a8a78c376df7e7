"""Wire-compatibility proof against the reference's OWN code (missing #3).

One federated job, two implementations: party *alice* runs the **unmodified reference**
(``baseline/_ref``; its generated stubs ``fed/grpc/pb4/fed_pb2_grpc.py`` — ``/GrpcService/SendData`` —
its ``GrpcSenderProxy`` / ``GrpcReceiverProxy`` and ``SendDataService``), party *bob* runs
``rayfed_b200`` with ``transport="grpc"``.  Both execute the same driver script; every cross-party
object crosses reference-stub → our receiver and our sender → reference servicer, pings included.

The reference needs ``ray`` (fed/api.py:23); where it is not installed the stdlib stand-in
``baseline/ray_shim`` supplies it (no rayfed_b200 code inside).
"""
import os
import subprocess
import sys
import textwrap

import pytest

from party_utils import make_addresses

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

if not os.path.isdir(os.path.join(REF, "fed")):
    pytest.skip("the reference is not installed under baseline/_ref", allow_module_level=True)

_SCRIPT = """
import json, os, sys
party, addresses, impl = sys.argv[1], json.loads(sys.argv[2]), sys.argv[3]
if impl == "reference":
    try:
        import ray
    except ImportError:
        sys.path.insert(0, os.path.join({root!r}, "baseline", "ray_shim"))
        import ray
    sys.path.insert(0, {ref!r})
    import fed
    assert os.path.abspath(fed.__file__).startswith({ref!r}), fed.__file__
    ray.init()
    comm = {{"grpc_retry_policy": {{"maxAttempts": 5, "initialBackoff": "0.2s", "maxBackoff": "1s",
                                   "backoffMultiplier": 2, "retryableStatusCodes": ["UNAVAILABLE"]}}}}
else:
    sys.path.insert(0, {root!r})
    import rayfed_b200 as fed
    ray = None
    comm = {{"transport": "grpc", "grpc_retry_policy": {{"maxAttempts": 5, "initialBackoff": "0.2s",
             "maxBackoff": "1s", "backoffMultiplier": 2, "retryableStatusCodes": ["UNAVAILABLE"]}}}}

fed.init(addresses=addresses, party=party, logging_level="warning",
         config={{"cross_silo_comm": comm, "barrier_on_initializing": True}})


@fed.remote
class MyActor:
    def __init__(self, value):
        self.value = value

    def inc(self, num):
        self.value = self.value + num
        return self.value

    def table(self, t):
        return {{k: v * 2 for k, v in t.items()}}


@fed.remote
def aggregate(val1, val2):
    return val1 + val2


@fed.remote
def blob(n):
    return bytes(range(256)) * n


actor_alice = MyActor.party("alice").remote(1)
actor_bob = MyActor.party("bob").remote(1)
val_alice = actor_alice.inc.remote(1)
val_bob = actor_bob.inc.remote(2)
total = aggregate.party("bob").remote(val_alice, val_bob)          # alice -> bob
assert fed.get(total) == 5                                          # bob -> alice (broadcast)
back = actor_alice.inc.remote(total)                                # bob -> alice as a task argument
assert fed.get(back) == 7
big = blob.party("alice").remote(4096)                              # 1 MiB alice -> bob
doubled = actor_bob.table.remote({{"n": 21}})
assert fed.get(doubled) == {{"n": 42}}
assert len(fed.get(big)) == 256 * 4096
for i in range(20):                                                 # seq-ids keep matching
    assert fed.get(aggregate.party("alice" if i % 2 else "bob").remote(actor_alice.inc.remote(0), i)) == 7 + i
fed.shutdown()
if ray is not None:
    ray.shutdown()
print("INTEROP-OK", party, impl, flush=True)
os._exit(0)
"""


@pytest.mark.parametrize("ref_party", ["alice", "bob"])
def test_reference_party_and_rayfed_b200_party_run_one_job(tmp_path, ref_party):
    addresses = make_addresses(["alice", "bob"])
    script = tmp_path / "interop.py"
    script.write_text(textwrap.dedent(_SCRIPT.format(root=ROOT, ref=REF)))
    import json

    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    procs = {}
    for party in ("alice", "bob"):
        impl = "reference" if party == ref_party else "ours"
        procs[party] = subprocess.Popen([sys.executable, str(script), party, json.dumps(addresses), impl],
                                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env,
                                        cwd=str(tmp_path), start_new_session=True)
    outs = {}
    for party, p in procs.items():
        try:
            outs[party] = p.communicate(timeout=120)[0]
        except subprocess.TimeoutExpired:
            import signal

            os.killpg(p.pid, signal.SIGKILL)
            outs[party] = "TIMEOUT\n" + (p.communicate()[0] or "")
    for party, p in procs.items():
        assert p.returncode == 0 and "INTEROP-OK" in outs[party], f"{party}:\n{outs[party][-3000:]}"
