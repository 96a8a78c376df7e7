"""Where one `many_tiny_tasks` iteration spends its time: both ends of an N-party run stamp
CLOCK_MONOTONIC (shared by all processes of a host) at the driver's call sites, inside the task
bodies and around the transport's send / broadcast / decode, and the medians of the differences
are printed as a timeline (party 0 = the aggregating party, party N-1 = a plain participant).
This is the measurement behind docs/enhancements/2026-09-host-path-wakeups.md.

    python tools/host_path_timeline.py 2        # or 4, 8
"""
import multiprocessing, sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ITERS = 2000

def run(party, parties, addresses, q):
    import rayfed_b200 as fed
    fed.init(addresses=addresses, party=party, logging_level="warning",
             config={"cross_silo_comm": {"transport": "native"}, "barrier_on_initializing": True})
    pc = time.perf_counter_ns
    ev = []
    from rayfed_b200.proxy.native import socket_proxy as SP
    def wrap(cls, name, label):
        orig = getattr(cls, name)
        def w(*a, **k):
            ev.append((label + "_in", pc()))
            try:
                return orig(*a, **k)
            finally:
                ev.append((label + "_out", pc()))
        setattr(cls, name, w)
    wrap(SP.NativeSenderProxy, "send", "send")
    wrap(SP.NativeSenderProxy, "broadcast", "bcast")
    wrap(SP._PayloadDecoder, "__call__", "decode")
    @fed.remote
    class MyActor:
        def run(self):
            ev.append(("run", pc())); return None
    @fed.remote
    class Aggregator:
        def aggr(self, *vals):
            ev.append(("aggr", pc())); return None
    actors = {p: MyActor.party(p).remote() for p in parties}
    aggregator = Aggregator.party(parties[0]).remote()
    for _ in range(200):
        fed.get(aggregator.aggr.remote(*[actors[p].run.remote() for p in parties]))
    ev.clear()
    for _ in range(ITERS):
        ev.append(("start", pc()))
        o = aggregator.aggr.remote(*[actors[p].run.remote() for p in parties])
        ev.append(("issued", pc()))
        fed.get(o)
        ev.append(("got", pc()))
    q.put((party, ev))
    fed.shutdown()

if __name__ == "__main__":
    parties = [f"p{i}" for i in range(N)]
    addresses = {p: f"127.0.0.1:{27100 + i}" for i, p in enumerate(parties)}
    ctx = multiprocessing.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=run, args=(p, parties, addresses, q)) for p in parties]
    [p.start() for p in procs]
    res = dict(q.get(timeout=300) for _ in parties)
    [p.join() for p in procs]
    import statistics as st
    def series(party, label): return [t for l, t in res[party] if l == label]
    p0, pl = parties[0], parties[-1]
    s0, i0, g0, a0, r0 = (series(p0, x) for x in ("start", "issued", "got", "aggr", "run"))
    sl, il, gl, rl = (series(pl, x) for x in ("start", "issued", "got", "run"))
    med = lambda xs: round(st.median(xs) / 1e3, 1)
    n = min(map(len, (s0, i0, g0, a0, r0, sl, il, gl, rl)))
    print("iter period p0      ", med([s0[k+1]-s0[k] for k in range(n-1)]))
    print("p0 issue (start->issued)", med([i0[k]-s0[k] for k in range(n)]))
    print("pL issue                ", med([il[k]-sl[k] for k in range(n)]))
    print("pL start->run body      ", med([rl[k]-sl[k] for k in range(n)]))
    print("pL run body -> p0 aggr  ", med([a0[k]-rl[k] for k in range(n)]))
    try:
        si, so = series(pl, "send_in"), series(pl, "send_out")
        d0i = series(p0, "decode_in"); d0o = series(p0, "decode_out")
        bi, bo = series(p0, "bcast_in"), series(p0, "bcast_out")
        dli, dlo = series(pl, "decode_in"), series(pl, "decode_out")
        per = len(d0i) // n
        print("counts", len(si), len(d0i), len(bi), len(dli), "decodes per iter at p0", per)
        print("  pL run -> send_in     ", med([si[k]-rl[k] for k in range(n)]))
        print("  pL send_in -> send_out", med([so[k]-si[k] for k in range(n)]))
        print("  pL send_out -> p0 last decode_in", med([d0i[k*per+per-1]-so[k] for k in range(n)]))
        print("  p0 decode dur (last)  ", med([d0o[k*per+per-1]-d0i[k*per+per-1] for k in range(n)]))
        print("  p0 last decode_out -> aggr", med([a0[k]-d0o[k*per+per-1] for k in range(n)]))
        print("  p0 aggr -> bcast_in   ", med([bi[k]-a0[k] for k in range(n)]))
        print("  p0 bcast dur          ", med([bo[k]-bi[k] for k in range(n)]))
        print("  p0 bcast_out -> pL decode_in", med([dli[k]-bo[k] for k in range(n)]))
        print("  pL decode dur         ", med([dlo[k]-dli[k] for k in range(n)]))
        print("  pL decode_out -> got  ", med([gl[k]-dlo[k] for k in range(n)]))
    except Exception as e:
        import traceback; traceback.print_exc()
    print("p0 start -> aggr        ", med([a0[k]-s0[k] for k in range(n)]))
    print("p0 aggr -> p0 got       ", med([g0[k]-a0[k] for k in range(n)]))
    print("p0 aggr -> pL got       ", med([gl[k]-a0[k] for k in range(n)]))
    print("p0 got -> pL got        ", med([gl[k]-g0[k] for k in range(n)]))
    print("pL got -> pL next start ", med([sl[k+1]-gl[k] for k in range(n-1)]))
