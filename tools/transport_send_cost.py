"""Cost of one native-transport send (call duration, send->ack round trip) as a function of the idle gap
between messages: shows what a sleeping ring reader costs (doorbell system call + wake-up).
"""
import sys, time, multiprocessing
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def receiver(addr, ev):
    from rayfed_b200.core import native
    n = native(); mb = n.Mailbox(); t = n.SocketTransport("bob", "job", mb)
    assert not t.listen(addr)
    ev.set(); time.sleep(30)
if __name__ == "__main__":
    ctx = multiprocessing.get_context("spawn")
    ev = ctx.Event()
    addr = "127.0.0.1:19555"
    p = ctx.Process(target=receiver, args=(addr, ev), daemon=True); p.start(); ev.wait(60)
    from rayfed_b200.core import native
    n = native(); t = n.SocketTransport("alice", "job", n.Mailbox())
    t.add_peer("bob", addr, 5, 0.1, 1.0, 2.0, 10.0, 0)
    payload = b"x" * 16
    t.send("bob", "0", "0", payload).result(10)
    for gap_us in (0, 100, 300, 1000):
        ds = []; rt = []
        for i in range(1, 1500):
            t0 = time.perf_counter_ns()
            a = t.send("bob", str(i), str(gap_us), payload)
            t1 = time.perf_counter_ns()
            a.result(5)
            t2 = time.perf_counter_ns()
            ds.append(t1 - t0); rt.append(t2 - t0)
            if gap_us:
                e = time.perf_counter_ns() + gap_us * 1000
                while time.perf_counter_ns() < e: pass
        ds.sort(); rt.sort()
        print(f"gap {gap_us:5d} us: send call median {ds[len(ds)//2]/1e3:.1f} us p90 {ds[len(ds)*9//10]/1e3:.1f}; send->ack median {rt[len(rt)//2]/1e3:.1f} us")
    t.stop(); p.terminate()
