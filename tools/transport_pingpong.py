"""Round-trip (send + ack) latency and large-message throughput of the native transport in one
process; run once with RAYFED_B200_SHM=0 and once with =1 to compare the socket and the
shared-memory data paths."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rayfed_b200.core import native
nat = native()
def bench(addr, n=20000, size=1):
    mb = nat.Mailbox()
    rx = nat.SocketTransport("bob", "job", mb)
    err = rx.listen(addr); assert not err, err
    tx = nat.SocketTransport("alice", "job", nat.Mailbox())
    tx.add_peer("bob", addr, 5, 0.05, 0.5, 2.0, 10.0)
    payload = b"x" * size
    for i in range(100):
        tx.send("bob", str(i), "w", payload, False, None).result(5.0); mb.claim(str(i), "w").result(5.0)
    t0 = time.perf_counter()
    for i in range(n):
        tx.send("bob", str(i), "d", payload, False, None).result(5.0)
        mb.claim(str(i), "d")   # pop the slot (value dropped)
    dt = time.perf_counter() - t0
    st = (tx.stats()["shm_conns"], rx.stats()["shm_conns"])
    tx.stop(); rx.stop()
    return dt / n * 1e6, st
for rep in range(3):
    r, st = bench("127.0.0.1:23998")
    r2, _ = bench("127.0.0.1:23997", n=200, size=4 << 20)
    print(f"shm={os.environ.get('RAYFED_B200_SHM','0')} rtt {r:.1f} us  shm_conns={st}   4MiB rtt {r2:.0f} us ({(4<<20)/r2/1e3:.2f} GB/s)")
