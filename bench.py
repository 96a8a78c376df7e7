#!/usr/bin/env python
"""bench.py — flagship benchmark of rayfed_b200 (driver contract).

Flagship step = one **FedAvg round of ResNet-50 bf16** through the public fed API: every party
(one per GPU) runs ``local_steps`` local SGD steps on its own synthetic 224x224 batches (H2D from
pinned host memory each step; fwd/bwd via PyTorch/cuDNN; *our* fused local-SGD kernel), then all
parties run *our* fused aggregate kernel (cross-party all-reduce + 1/N + all-gather over NVLink
inside one kernel), and the round's losses are fetched with the collective ``fed.get``.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Rank 0 prints ONE JSON line.  ``value`` = samples/s summed over all parties (device-timed with
CUDA events, max over parties); ``e2e`` = the same through the public API with wall-clock around K
rounds including the per-round ``fed.get`` of the losses.  Secondary keys:

``aggregate``        K2 alone: µs, fraction of the NVLink/HBM roofline (NVLS byte count when the
                     multicast path is in use; link peak measured in-run by the push sweep)
``agg_check``        N>1: after an aggregate every party's weights are bit-identical and equal the
                     fp32 oracle ``stack(weights).float().mean(0)`` computed over the DAG at p0
``push``             N>1: the cross-party tensor push *through the API*
                     (``sink.party(b).remote(make.party(a).remote(n))``), 1 KiB … 1 GiB: wall-clock
                     per call, pipelined GB/s, and the device-timed push kernel
``many_tiny_tasks``  the reference's own benchmark (N-party generalisation)
``comm_bound``       rounds/s with ONE local step per round and aggregate-only rounds/s

``--impl reference`` runs the unmodified reference from ``baseline/_ref`` through its own public API
(``baseline/ref_arm.py``; Ray if importable, otherwise the ``baseline/ray_shim`` stand-in, labelled).
``--impl reference-protocol`` measures the same FedAvg round with the aggregation done the
reference way on OUR runtime over the wire-compatible gRPC transport.
``--config llama-lora`` runs BASELINE config #4 (Llama-3-8B-shaped federated LoRA, fused Adam).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 32
LOCAL_STEPS = 4
IMG = 224
NUM_CLASSES = 1000
NVLINK_REF_GBPS = 770.0  # peer-copy reference of the profiling recipe (900 nominal)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (rank 0)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(smax) if smax else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ---------------------------------------------------------------------------
def make_worker_cls():
    import torch

    from rayfed_b200.fl import FedAvgWorker
    from rayfed_b200.ops import kernels as K

    class BenchWorker(FedAvgWorker):
        """FedAvgWorker + party-local synthetic dataset in pinned host memory + timers."""

        def __init__(self, parties, party, seed=0, use_cuda_graph=True, kind="fedavg", fused=True):
            from rayfed_b200.models import resnet50

            super().__init__(lambda: resnet50(NUM_CLASSES), parties=parties, party=party, lr=0.05, momentum=0.9,
                             weight_decay=5e-5, kind=kind, seed=seed, use_cuda_graph=use_cuda_graph,
                             fused_aggregate=fused)
            g = torch.Generator().manual_seed(1000 + list(parties).index(party))
            nbuf = 8  # ring of distinct pinned batches (77 MB) — larger than any cache on the path
            self._x = [torch.randn(BATCH, 3, IMG, IMG, generator=g).to(torch.bfloat16).pin_memory()
                       for _ in range(nbuf)]
            self._y = [torch.randint(0, NUM_CLASSES, (BATCH,), generator=g).pin_memory() for _ in range(nbuf)]
            self._i = 0
            self.batch_fn = self.next_batch
            self.h2d_bytes_per_step = self._x[0].numel() * 2 + self._y[0].numel() * 8
            self._ev = {}

        def next_batch(self, r, i):
            j = self._i % len(self._x)
            self._i += 1
            return self._x[j], self._y[j]

        def round(self, steps):
            """L local steps (H2D of each batch from pinned memory) + fused aggregate."""
            loss = self.local_round(steps)
            self.agg.aggregate()
            return loss

        def mark(self, name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream())
            self._ev[name] = ev
            return True

        def elapsed_ms(self, a, b):
            self._ev[b].synchronize()
            return self._ev[a].elapsed_time(self._ev[b])

        def launches(self):
            return dict(K.launch_counter.by_kernel), K.launch_counter.count

        def reset_launches(self):
            K.launch_counter.reset()
            return True

        def sync(self):
            torch.cuda.synchronize()
            return True

        def info(self):
            return {"bucket_numel": self.bucket.numel,
                    "multicast": self.agg.uses_multicast if self.agg is not None else False,
                    "backend": self.bucket.buffer.backend if self.bucket.buffer is not None else "local",
                    "h2d_bytes_per_step": self.h2d_bytes_per_step}

        def time_aggregate(self, iters=20):
            """K2 alone, device-timed (collective: every party calls it)."""
            s = torch.cuda.current_stream()
            for _ in range(3):
                self.agg.aggregate()
            a = torch.cuda.Event(enable_timing=True)
            b = torch.cuda.Event(enable_timing=True)
            a.record(s)
            for _ in range(iters):
                self.agg.aggregate()
            b.record(s)
            b.synchronize()
            return a.elapsed_time(b) / iters * 1e3  # µs

        def agg_phases(self):
            return self.agg.phase_breakdown(5)

        def time_aggregate_variants(self, iters=20):
            """The same aggregate with the other kernel / the other link path (collective)."""
            from rayfed_b200.fl.fedavg import FedAggregator

            def timed(agg):
                s = torch.cuda.current_stream()
                for _ in range(3):
                    agg.aggregate()
                a = torch.cuda.Event(enable_timing=True)
                b = torch.cuda.Event(enable_timing=True)
                a.record(s)
                for _ in range(iters):
                    agg.aggregate()
                b.record(s)
                b.synchronize()
                return a.elapsed_time(b) / iters * 1e3

            out = {}
            mc_now = self.agg.uses_multicast
            self.agg._plan.set_force_generic(True)
            out["generic_kernel_" + ("nvls" if mc_now else "p2p")] = timed(self.agg)
            self.agg._plan.set_force_generic(False)
            if self.group.size > 1:
                other = FedAggregator(self.bucket, self.group, self.channel, kind="fedavg", heap=self.heap,
                                      use_multicast=not mc_now)
                if other.uses_multicast != mc_now:
                    tag = "nvls" if other.uses_multicast else "p2p"
                    out["pipelined_kernel_" + tag] = timed(other)
                    other._plan.set_force_generic(True)
                    out["generic_kernel_" + tag] = timed(other)
                    other._plan.set_force_generic(False)
                # grid sweep of the pipelined kernel on both link paths (same cap in every party)
                for agg, tag in ((self.agg, "nvls" if mc_now else "p2p"), (other, "p2p" if mc_now else "nvls")):
                    if (tag == "nvls") != agg.uses_multicast:
                        continue
                    for cap in (16, 32, 64, 148):
                        agg._plan.set_grid_cap(cap)
                        out[f"pipelined_kernel_{tag}_grid{cap}"] = timed(agg, ) if True else None
                    agg._plan.set_grid_cap(0)
            return out

        def flat_weights_host(self):
            return self.bucket.flat_w.cpu()

        def load_flat_weights(self, w):
            self.bucket.flat_w.copy_(w.to(self.bucket.flat_w.device))
            return True

        # -- agg_check: perturb (so the parties really differ), snapshot, aggregate, snapshot ----
        def perturb_and_snapshot(self, rank):
            g = torch.Generator(device=self.bucket.flat_w.device).manual_seed(77 + rank)
            noise = torch.randn(self.bucket.numel, device=self.bucket.flat_w.device, generator=g) * 1e-2
            self.bucket.flat_w.add_(noise.to(self.bucket.flat_w.dtype))
            return self.bucket.flat_w.clone()

        def aggregate_and_snapshot(self):
            self.agg.aggregate()
            return self.bucket.flat_w.clone()

    return BenchWorker


def addresses_for(n, base_port):
    return {f"p{i}": f"127.0.0.1:{base_port + i}" for i in range(n)}


def _median(xs):
    return statistics.median(xs) if xs else None


def push_block(fed, parties, rank, quick=False):
    """BASELINE config #5 through the public API: p0 produces a CUDA tensor, p1 consumes it."""
    import torch

    from rayfed_b200.proxy.nvlink import nvlink_proxy

    a, b = parties[0], parties[1]
    ch = nvlink_proxy.current_channel()

    @fed.remote
    class Peer:
        def __init__(self):
            self._cache = {}

        def make(self, nbytes):
            t = self._cache.get(nbytes)
            if t is None:
                t = self._cache[nbytes] = torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device="cuda")
            return t

        def none(self):
            return None

        def sink(self, t):
            # touches both ends on the device and reads 2 bytes back: the data really arrived
            return 0 if t is None else int(t[0]) + int(t[-1])

        def start_timing(self):
            c = nvlink_proxy.current_channel()
            if c is not None:
                c.push_times = []
            return True

        def stop_timing(self):
            c = nvlink_proxy.current_channel()
            if c is None or c.push_times is None:
                return []
            torch.cuda.synchronize()
            out = [(nb, e0.elapsed_time(e1) * 1e3) for nb, e0, e1 in c.push_times]
            c.push_times = None
            return out

        def stats(self):
            c = nvlink_proxy.current_channel()
            return dict(c.stats) if c is not None else {}

    pa = Peer.party(a).remote()
    pb = Peer.party(b).remote()

    def one(n):
        src = pa.make.remote(n) if n else pa.none.remote()
        return fed.get(pb.sink.remote(src))

    sizes = [1 << 10, 1 << 16, 1 << 20, 1 << 24, 1 << 26, 1 << 28, 1 << 30]
    if quick:
        sizes = [1 << 10, 1 << 20, 1 << 26]
    for _ in range(5):
        one(0)
    base_us = []
    for _ in range(30):
        t0 = time.perf_counter()
        one(0)
        base_us.append((time.perf_counter() - t0) * 1e6)
    base = _median(base_us)
    rows = []
    for n in sizes:
        it = 20 if n <= (1 << 24) else (8 if n <= (1 << 28) else 4)
        for _ in range(3):
            one(n)
        fed.get(pa.start_timing.remote())
        lat = []
        for _ in range(it):
            t0 = time.perf_counter()
            one(n)
            lat.append((time.perf_counter() - t0) * 1e6)
        # pipelined: M pushes in flight, one fed.get at the end
        m = it
        t0 = time.perf_counter()
        outs = [pb.sink.remote(pa.make.remote(n)) for _ in range(m)]
        fed.get(outs)
        piped_s = time.perf_counter() - t0
        ktimes = fed.get(pa.stop_timing.remote())
        kus = _median([us for nb, us in ktimes if nb >= n]) if ktimes else None
        # chunked tensors: several messages per tensor; sum the kernel time per tensor instead
        if ktimes and kus is None:
            per_tensor = sum(us for _, us in ktimes) / max(1, (it + m))
            kus = per_tensor
        rows.append({"bytes": n, "api_us": _median(lat), "api_us_minus_empty_call": _median(lat) - base,
                     "pipelined_GBps": n * m / piped_s / 1e9,
                     "kernel_us": kus, "kernel_GBps": (n / (kus * 1e-6) / 1e9) if kus else None,
                     "messages_per_tensor": max(1, len(ktimes) // max(1, it + m)) if ktimes else None})
    # K1b: fed.get broadcast (owner -> all) with and without the multicast push
    bcast = []
    if ch is not None:
        for n in ([1 << 20, 1 << 26] if not quick else [1 << 20]):
            row = {"bytes": n, "n_receivers": len(parties) - 1}
            for mode in ("multicast", "unicast"):
                nvlink_proxy._MC_PUSH_ENABLED = mode == "multicast"
                for _ in range(2):
                    fed.get(pa.make.remote(n))
                t0 = time.perf_counter()
                k = 6
                for _ in range(k):
                    fed.get(pa.make.remote(n))
                row[mode + "_us"] = (time.perf_counter() - t0) / k * 1e6
            nvlink_proxy._MC_PUSH_ENABLED = True
            row["multicast_available"] = bool(ch.can_multicast)
            bcast.append(row)
    stats = fed.get([pa.stats.remote(), pb.stats.remote()])
    peak = max([r["kernel_GBps"] for r in rows if r["kernel_GBps"]] or [0.0])
    return {"path": f"{a}->{b} through fed.remote/fed.get (sink(make(n)))", "empty_call_us": base, "sizes": rows,
            "broadcast": bcast, "kernel_peak_GBps": peak,
            "zero_copy_tensors": stats[1].get("zero_copy_tensors"), "copied_tensors": stats[1].get("copied_tensors"),
            "note": "api_us: wall clock of one produce->push->consume->fed.get chain; kernel_*: CUDA events around "
                    "the push kernel on the sender's push stream; tensors > arena/4 are chunked (messages_per_tensor)"}


def run_ours(args):
    import torch

    import rayfed_b200 as fed
    from rayfed_b200.ops import kernels as K  # noqa: F401

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    n = args.gpus
    assert world == n, f"--gpus {n} but WORLD_SIZE={world}; launch with torchrun --nproc-per-node {n}"
    torch.cuda.set_device(local_rank)
    base = env_int("MASTER_PORT", 29500) + 200
    parties = [f"p{i}" for i in range(n)]
    me = parties[rank]
    transport = "nvlink" if args.impl == "ours" else "grpc"
    comm = {"transport": transport, "retry_max_attempts": 200, "retry_initial_backoff_s": 0.05,
            "retry_max_backoff_s": 0.5, "retry_backoff_multiplier": 1.5, "timeout_in_ms": 120000,
            "device_arena_bytes": args.arena_mib << 20,
            "messages_max_size_in_bytes": 1 << 30}
    fed.init(addresses=addresses_for(n, base), party=me, logging_level="warning",
             config={"cross_silo_comm": comm, "device": f"cuda:{local_rank}", "barrier_on_initializing": True},
             job_name=f"bench-{base}")

    Worker = fed.remote(make_worker_cls())
    workers = {p: Worker.party(p).remote(parties, p, use_cuda_graph=not args.no_graph, fused=args.impl == "ours")
               for p in parties}
    info = fed.get(workers[parties[0]].info.remote())
    L, K_steps, W = args.local_steps, args.steps, args.warmup

    @fed.remote
    def mean_weights(*ws):
        return torch.stack([w.float() for w in ws]).mean(0).to(ws[0].dtype)

    def submit_round(steps=None):
        """Issue one FedAvg round on every party; returns the FedObjects of the losses."""
        steps = L if steps is None else steps
        if args.impl == "ours":
            return [workers[p].round.remote(steps) for p in parties]
        # reference-style aggregation through the DAG API over the gRPC transport:
        # local steps, weights -> p0, mean, -> every party
        losses = [workers[p].local_round.remote(steps) for p in parties]
        ws = [workers[p].flat_weights_host.remote() for p in parties]
        m = mean_weights.party(parties[0]).remote(*ws)
        for p in parties:
            workers[p].load_flat_weights.remote(m)
        return losses

    def one_round(fetch_loss):
        losses = submit_round()
        if fetch_loss:
            return [float(x) for x in fed.get(losses)]
        return None

    # ---------------- warm-up (includes CUDA-graph capture, cuDNN autotune) ----------------
    for _ in range(max(W, 3)):
        one_round(True)
    fed.get([workers[p].sync.remote() for p in parties])
    fed.get([workers[p].reset_launches.remote() for p in parties])

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---------------- device-timed region: K rounds, CUDA events on each party's stream ----
    fed.get([workers[p].sync.remote() for p in parties])
    for p in parties:
        workers[p].mark.remote("t0")
    for _ in range(K_steps):
        one_round(False)
    for p in parties:
        workers[p].mark.remote("t1")
    ms = fed.get([workers[p].elapsed_ms.remote("t0", "t1") for p in parties])
    dev_ms_per_round = max(ms) / K_steps
    launches = fed.get([workers[p].launches.remote() for p in parties])
    gpu_launches = sum(c for _, c in launches)

    # ---------------- end-to-end through the public API: per-round fed.get of the losses -----
    fed.get([workers[p].sync.remote() for p in parties])
    t0 = time.perf_counter()
    last = None
    prev = None
    for _ in range(K_steps):
        # the controller issues round r, then reads round r-1's losses (D2H) while r runs:
        # every step's result is still fetched inside the timed region
        cur = submit_round()
        if prev is not None:
            last = [float(x) for x in fed.get(prev)]
        prev = cur
    last = [float(x) for x in fed.get(prev)]
    fed.get([workers[p].sync.remote() for p in parties])
    e2e_s = time.perf_counter() - t0

    # max over parties of the wall-clock (every party ran the same loop; exchange the numbers)
    @fed.remote
    def _ident(x):
        return x

    e2e_all = fed.get([_ident.party(p).remote(e2e_s) for p in parties]) if n > 1 else [e2e_s]
    e2e_s = max(e2e_all)
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- secondary metrics -------------------------------------------------------
    samples_per_round = n * L * BATCH
    value = samples_per_round / (dev_ms_per_round * 1e-3)
    e2e_value = samples_per_round * K_steps / e2e_s
    out = {
        "metric": "fedavg_resnet50_bf16_samples_per_sec",
        "value": value,
        "unit": "samples/s",
        "n_gpus": n,
        "steps": K_steps,
        "warmup": max(W, 3),
        "ms_per_step": dev_ms_per_round,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "vs_baseline_note": "BASELINE.md publishes no number; the driver's --impl reference arm (unmodified "
                            "reference over baseline/ray_shim) and --impl reference-protocol are the comparisons",
        "dtype": "bf16",
        "data": "synthetic (random 224x224 images in pinned host memory, random-init ResNet-50)",
        "impl": args.impl,
        "rounds_per_sec": 1e3 / dev_ms_per_round,
        "config": {
            "model": "ResNet-50 (25.56M params)",
            "global_batch": n * BATCH,
            "seq_len": None,
            "image": IMG,
            "local_steps_per_round": L,
            "per_party_batch": BATCH,
            "parallelism": f"fedavg-dp{n} (one party per GPU)",
            "step_definition": "1 FedAvg round = local_steps x (H2D batch + fwd/bwd + fused local SGD) + fused aggregate",
            "l2_policy": "inputs larger than L2: weights+grads+momentum 256 MB and >1 GB activations per step; "
                         "8 distinct pinned batches rotate",
            "transport": transport,
            "aggregate_backend": info,
            "cuda_graph": not args.no_graph,
        },
        "e2e": {
            "value": e2e_value,
            "unit": "samples/s",
            "rounds_per_sec": K_steps / e2e_s,
            "h2d_bytes_per_step": info["h2d_bytes_per_step"] * L,
            "d2h_bytes_per_step": 4 * n,
            "last_losses": last,
        },
        "gpu_launches": gpu_launches,
        "gpu_launches_by_kernel": launches[0][0],
        "clocks": clocks,
    }

    # The headline is measured; everything below is secondary.  A secondary block may fail (caught
    # by `guarded`) — and it must not be able to HANG the line away either: past the deadline every
    # rank leaves, rank 0 printing what it has.
    secondary_deadline_s = float(os.environ.get("RAYFED_B200_BENCH_SECONDARY_DEADLINE_S", "420"))
    current_block = ["-"]

    def bail():
        import copy

        for _ in range(5):
            try:
                snap = copy.deepcopy(out)
                break
            except RuntimeError:  # the main thread was adding a key
                time.sleep(0.01)
        else:
            snap = {k: out[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step") if k in out}
        snap["watchdog"] = (f"secondary blocks exceeded {secondary_deadline_s:.0f} s in '{current_block[0]}': "
                            "partial result, the headline and e2e numbers are complete")
        if rank == 0:
            print(json.dumps(snap), flush=True)
        os._exit(0)

    watchdog = threading.Timer(secondary_deadline_s, bail)
    watchdog.daemon = True
    watchdog.start()

    def guarded(name, fn):
        current_block[0] = name
        try:
            out[name] = fn()
        except Exception as e:  # noqa: BLE001  (a secondary metric must never cost the headline)
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}

    if args.impl == "ours":
        # comm-bound variants of the headline: 1 local step per round, and the aggregate alone
        def comm_bound():
            for _ in range(2):
                fed.get(submit_round(1))
            fed.get([workers[p].sync.remote() for p in parties])
            for p in parties:
                workers[p].mark.remote("c0")
            k = max(5, K_steps)
            for _ in range(k):
                submit_round(1)
            for p in parties:
                workers[p].mark.remote("c1")
            ms1 = max(fed.get([workers[p].elapsed_ms.remote("c0", "c1") for p in parties])) / k
            return {"local_steps_1_rounds_per_sec": 1e3 / ms1, "local_steps_1_ms_per_round": ms1}

        guarded("comm_bound", comm_bound)

        def aggregate_block():
            # every party's earlier device work (the push sweep's last transfers) is complete before the
            # collective is timed
            fed.get([workers[p].sync.remote() for p in parties])
            agg_us = max(fed.get([workers[p].time_aggregate.remote(20) for p in parties]))
            nbytes = info["bucket_numel"] * 2
            peaks = measured_peaks()
            hbm = float(peaks.get("hbm_gbs", 6485.0)) * 1e9
            link = NVLINK_REF_GBPS * 1e9
            push = out.get("push") if isinstance(out.get("push"), dict) else None
            link_src = f"{NVLINK_REF_GBPS:.0f} GB/s/direction peer-copy reference (profiling recipe)"
            if info["multicast"]:
                # NVLS: the switch reduces/replicates; per direction a GPU moves its whole buffer once:
                # serve (N-1)/N*S to the reducers + S/N multimem.st out; receive S/N reduced + (N-1)/N*S
                link_bytes = float(nbytes) if n > 1 else 0.0
                link_desc = "NVLS: ~S bytes per direction per GPU"
            else:
                link_bytes = 2.0 * (n - 1) / n * nbytes if n > 1 else 0.0
                link_desc = "P2P two-shot: 2(N-1)/N*S bytes per direction per GPU"
            # local HBM: every byte of my buffer is read once (by me or a peer) and written once
            roof_s = max(link_bytes / link, 2.0 * nbytes / hbm)
            d = {"us": agg_us, "bytes": nbytes, "roofline_us": roof_s * 1e6,
                 "fraction_of_roofline": (roof_s * 1e6) / agg_us if agg_us else None,
                 "aggregate_only_rounds_per_sec": 1e6 / agg_us if agg_us else None,
                 "multicast": bool(info["multicast"]),
                 "roofline": f"max({link_desc} / {link_src}, 2S / {hbm / 1e9:.0f} GB/s measured HBM)"}
            try:
                var = fed.get([workers[p].time_aggregate_variants.remote() for p in parties])
                d["variants_us"] = {k: max(v[k] for v in var) for k in var[0]}
                d["variants_us"]["pipelined_kernel_" + ("nvls" if info["multicast"] else "p2p")] = agg_us
            except Exception as e:  # noqa: BLE001
                d["variants_error"] = repr(e)[:200]
            try:
                ph = fed.get([workers[p].agg_phases.remote() for p in parties])
                d["phases_us_party0"] = ph[0]
                if all(ph):
                    d["phases_us_max_over_parties"] = {k: max(x[k] for x in ph) for k in ph[0] if k != "ctas"}
            except Exception as e:  # noqa: BLE001
                d["phases_error"] = repr(e)[:200]
            if push and push.get("kernel_peak_GBps"):
                d["nvlink_peer_GBps_measured_in_run"] = push["kernel_peak_GBps"]
                d["fraction_of_roofline_vs_in_run_link_peak"] = (
                    max(link_bytes / (push["kernel_peak_GBps"] * 1e9), 0.0) * 1e6 / agg_us if agg_us and n > 1 else None)
            return d

        if n > 1 and not args.no_push:
            guarded("push", lambda: push_block(fed, parties, rank, quick=args.quick))
        guarded("aggregate", aggregate_block)

        if n > 1:
            def agg_check():
                before = [workers[p].perturb_and_snapshot.remote(i) for i, p in enumerate(parties)]
                after = [workers[p].aggregate_and_snapshot.remote() for p in parties]

                @fed.remote
                def oracle(nb, *snaps):
                    bs, as_ = snaps[:nb], snaps[nb:]
                    exp = torch.stack([b.float() for b in bs]).mean(0)
                    a0 = as_[0].float()
                    max_err = float((a0 - exp).abs().max())
                    # one bf16 ulp (8 significant bits) at the largest magnitude: 2^(floor(log2 max) - 7)
                    import math

                    ulp = 2.0 ** (math.floor(math.log2(max(float(exp.abs().max()), 1e-30))) - 7)
                    equal = all(bool(torch.equal(as_[0], x)) for x in as_[1:])
                    return {"parties_equal": equal, "max_err": max_err, "bf16_ulp_at_max": ulp,
                            "within_1ulp": max_err <= ulp, "n_params": int(exp.numel())}

                return fed.get(oracle.party(parties[0]).remote(len(parties), *before, *after))

            guarded("agg_check", agg_check)
    guarded("many_tiny_tasks", lambda: many_tiny_tasks(fed, parties, iters=args.tiny_iters))
    current_block[0] = "shutdown"
    fed.shutdown()
    watchdog.cancel()
    if rank == 0:
        print(json.dumps(out), flush=True)


def many_tiny_tasks(fed, parties, iters=2000):
    """The reference's benchmark (benchmarks/many_tiny_tasks_benchmark.py:23-62),
    N-party generalisation: per iteration one no-op actor call per party, one
    aggregator call consuming all of them, one blocking fed.get."""

    @fed.remote
    class Tiny:
        def f(self):
            return None

        def agg(self, *xs):
            return None

    actors = {p: Tiny.party(p).remote() for p in parties}
    for _ in range(50):
        fed.get(actors[parties[0]].agg.remote(*[actors[p].f.remote() for p in parties]))
    t0 = time.perf_counter()
    for _ in range(iters):
        fed.get(actors[parties[0]].agg.remote(*[actors[p].f.remote() for p in parties]))
    dt = time.perf_counter() - t0
    n = len(parties)
    try:
        from rayfed_b200.core import native

        spin = native().spin_info()
    except Exception:  # noqa: BLE001
        spin = None
    return {
        "iterations": iters,
        "spin_policy": spin,  # what fed.init picked from usable CPUs and parties per host
        "per_iteration_ms": dt / iters * 1e3,
        "tasks_per_sec": (n + 1) * iters / dt,
        "tasks_per_iteration": n + 1,
        "cross_party_messages_per_iteration": 2 * (n - 1),
    }


def run_reference(args):
    """The unmodified reference through its own public API (baseline/ref_arm.py)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import ref_arm
    except Exception as e:  # noqa: BLE001
        if env_int("RANK", 0) == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"baseline/ref_arm.py failed to import: {e!r}"}),
                  flush=True)
        return
    ref_arm.run(args, clock_sampler_cls=ClockSampler)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-protocol"])
    ap.add_argument("--config", default="resnet50", choices=["resnet50", "llama-lora"])
    ap.add_argument("--local-steps", type=int, default=LOCAL_STEPS)
    ap.add_argument("--tiny-iters", type=int, default=2000)
    ap.add_argument("--arena-mib", type=int, default=1024, help="device arena region per (src,dst) pair")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-push", action="store_true", help="skip the API-level push sweep")
    ap.add_argument("--quick", action="store_true", help="shorter secondary sweeps")
    args = ap.parse_args()
    if args.config == "llama-lora":
        sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
        import fedavg_llama_lora

        return fedavg_llama_lora.bench_main(args)
    if args.impl == "reference":
        return run_reference(args)
    run_ours(args)


if __name__ == "__main__":
    main()
