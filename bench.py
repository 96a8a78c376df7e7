#!/usr/bin/env python
"""bench.py — flagship benchmark of rayfed_b200 (driver contract).

Flagship step = one **FedAvg round of ResNet-50 bf16** through the public fed
API: every party (one per GPU) runs ``local_steps`` local SGD steps on its own
synthetic 224x224 batches (H2D from pinned host memory each step; fwd/bwd via
PyTorch/cuDNN; *our* fused local-SGD kernel), then all parties run *our* fused
aggregate kernel (cross-party all-reduce + 1/N + all-gather over NVLink inside
one kernel), and the round's losses are fetched with the collective ``fed.get``.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Rank 0 prints ONE JSON line.  ``value`` = samples/s summed over all parties
(device-timed with CUDA events, max over parties); ``e2e`` = the same through
the public API with wall-clock around K rounds including the per-round
``fed.get`` of the losses.  Secondary keys: ``rounds_per_sec``,
``many_tiny_tasks`` (the reference's own benchmark, N-party generalisation),
``push`` (K1 GB/s party0→party1) and ``aggregate`` (K2 µs + roofline fraction).

``--impl reference`` runs the unmodified reference from ``baseline/_ref`` if it
is importable (it needs Ray, which is not in the offline wheelhouse) and
otherwise prints ``{"impl": "reference", "unavailable": ...}``.
``--impl reference-protocol`` measures the same FedAvg round with the
aggregation done the reference way (weights pushed to one party over the
wire-compatible gRPC transport, mean, pushed back).
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


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


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

        def flat_weights_host(self):
            return self.bucket.flat_w.cpu()

        def load_flat_weights(self, w):
            self.bucket.flat_w.copy_(w.to(self.bucket.flat_w.device))
            return True

    return BenchWorker


def addresses_for(n, base_port):
    return {f"p{i}": f"127.0.0.1:{base_port + i}" for i in range(n)}


def run_ours(args):
    import torch

    import rayfed_b200 as fed
    from rayfed_b200.ops import kernels as K

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
            "device_arena_bytes": 128 << 20,
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

    def submit_round():
        """Issue one FedAvg round on every party; returns the FedObjects of the losses."""
        if args.impl == "ours":
            return [workers[p].round.remote(L) for p in parties]
        # reference-style aggregation through the DAG API over the gRPC transport:
        # local steps, weights -> p0, mean, -> every party
        losses = [workers[p].local_round.remote(L) for p in parties]
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
    agg_us = None
    if args.impl == "ours":
        agg_us = max(fed.get([workers[p].time_aggregate.remote(20) for p in parties]))
    tiny = many_tiny_tasks(fed, parties, iters=args.tiny_iters)

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
        "many_tiny_tasks": tiny,
    }
    if agg_us is not None:
        nbytes = info["bucket_numel"] * 2
        # bytes that must leave each GPU: serve (N-1)/N of my buffer to readers + write (N-1)/N to peers
        link_bytes = 2 * (n - 1) / n * nbytes if n > 1 else 0
        hbm_bytes = 2 * nbytes / max(n, 1) * (n if n > 1 else 1) + 2 * nbytes / max(n, 1)
        roof_s = max(link_bytes / 770e9, (2 * nbytes / n) / 6485e9)
        out["aggregate"] = {"us": agg_us, "bytes": nbytes, "roofline_us": roof_s * 1e6,
                            "fraction_of_roofline": (roof_s * 1e6) / agg_us if agg_us else None,
                            "roofline": "max(2*(N-1)/N*bytes / 770 GB/s measured NVLink, local shard r+w / 6485 GB/s measured HBM)"}
    fed.shutdown()
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
    return {
        "iterations": iters,
        "per_iteration_ms": dt / iters * 1e3,
        "tasks_per_sec": (n + 1) * iters / dt,
        "tasks_per_iteration": n + 1,
        "cross_party_messages_per_iteration": 2 * (n - 1),
    }


def run_reference(args):
    """The unmodified reference needs Ray; say so if it is not importable."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    why = None
    if not os.path.isdir(os.path.join(ref_dir, "fed")):
        why = "reference not installed under baseline/_ref"
    else:
        try:
            import ray  # noqa: F401
        except Exception as e:  # noqa: BLE001
            why = (f"reference is installed (pip --no-deps) but cannot import: it requires Ray "
                   f"(fed/api.py:23 `import ray`) and no ray wheel exists in /opt/wheelhouse: {e!r}")
    if why is None:
        sys.path.insert(0, ref_dir)
        try:
            import fed  # noqa: F401

            why = "reference import succeeded but no GPU FedAvg entry point exists in the reference to run"
        except Exception as e:  # noqa: BLE001
            why = f"reference import failed: {e!r}"
    if env_int("RANK", 0) == 0:
        print(json.dumps({"impl": "reference", "unavailable": why}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-protocol"])
    ap.add_argument("--local-steps", type=int, default=LOCAL_STEPS)
    ap.add_argument("--tiny-iters", type=int, default=2000)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    run_ours(args)


if __name__ == "__main__":
    main()
