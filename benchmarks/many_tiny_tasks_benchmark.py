"""The reference's only benchmark, generalised to N parties.

Reference: benchmarks/many_tiny_tasks_benchmark.py:23-62 — two parties, 10 000
iterations of {one no-op actor call per party, one aggregator call, a blocking
fed.get}; it prints "per task overhead (ms)".  Here every party is one process
of this script (spawned), the transport is selectable, and the result is also
reported as tasks/sec.

    python benchmarks/many_tiny_tasks_benchmark.py --parties 2 --iters 10000 --transport native
    python benchmarks/many_tiny_tasks_benchmark.py --parties 2 --iters 2000 --transport grpc   # reference protocol
"""
import argparse
import json
import multiprocessing
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(party, parties, addresses, transport, iters, q):
    import rayfed_b200 as fed

    fed.init(addresses=addresses, party=party, logging_level="warning",
             config={"cross_silo_comm": {"transport": transport, "retry_initial_backoff_s": 0.05,
                                         "retry_max_attempts": 100, "retry_max_backoff_s": 0.5},
                     "barrier_on_initializing": True})

    @fed.remote
    class MyActor:
        def run(self):
            return None

    @fed.remote
    class Aggregator:
        def aggr(self, *vals):
            return None

    actors = {p: MyActor.party(p).remote() for p in parties}
    aggregator = Aggregator.party(parties[0]).remote()
    for _ in range(100):
        fed.get(aggregator.aggr.remote(*[actors[p].run.remote() for p in parties]))
    start = time.perf_counter()
    for _ in range(iters):
        fed.get(aggregator.aggr.remote(*[actors[p].run.remote() for p in parties]))
    dt = time.perf_counter() - start
    if party == parties[0]:
        n = len(parties)
        q.put({
            "parties": n, "transport": transport, "iterations": iters,
            "per_iteration_ms": dt / iters * 1e3,
            "per_task_overhead_ms": dt / iters * 1e3,  # the reference's printed quantity
            "tasks_per_sec": (n + 1) * iters / dt,
        })
    fed.shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--parties", type=int, default=2)
    ap.add_argument("--iters", type=int, default=10000)
    ap.add_argument("--transport", default="native", choices=["native", "grpc"])
    ap.add_argument("--base-port", type=int, default=13100)
    args = ap.parse_args()
    parties = [f"p{i}" for i in range(args.parties)]
    addresses = {p: f"127.0.0.1:{args.base_port + i}" for i, p in enumerate(parties)}
    ctx = multiprocessing.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=run, args=(p, parties, addresses, args.transport, args.iters, q)) for p in parties]
    [p.start() for p in procs]
    [p.join() for p in procs]
    print(json.dumps(q.get(timeout=5)))


if __name__ == "__main__":
    main()
