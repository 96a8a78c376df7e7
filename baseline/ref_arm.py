"""Reference arm of bench.py: the UNMODIFIED reference (ray-project/rayfed, installed under
``baseline/_ref``) running the same metric and config as the product arm through its own
public API and stock code path — ``fed.init / fed.remote / fed.get`` over its gRPC proxies.

Nothing from ``rayfed_b200`` is imported here: the model is a plain ``torch.nn`` ResNet-50,
the optimizer ``torch.optim.SGD``, the aggregation is the reference's own idiom (weights of
every party → one party, mean, → every party; fed/tests/test_fed_get.py:45-78), and the tiny-task
loop is the reference's benchmark (benchmarks/many_tiny_tasks_benchmark.py:23-62) generalised to
N parties exactly like the product arm's.

Ray: the reference imports it (fed/api.py:23) and no wheel exists offline.  When the real ``ray``
is importable it is used; otherwise ``baseline/ray_shim`` (a stdlib stand-in for the missing
dependency, see its README) is put on ``sys.path`` and the output line says
``"substrate": "ray-shim"`` — a *lower bound* on the reference's cost (actor calls are function
calls instead of process hops).
"""
from __future__ import annotations

import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
BATCH = 32
IMG = 224
NUM_CLASSES = 1000


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def import_reference():
    """(fed module, ray module, substrate) or raises ImportError with a one-line reason."""
    ref_dir = os.path.join(HERE, "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "fed")):
        raise ImportError("reference not installed under baseline/_ref (pip install --target baseline/_ref "
                          "/root/reference)")
    substrate = "ray"
    try:
        import ray  # noqa: F401
        if getattr(ray, "__shim__", False):
            substrate = "ray-shim"
    except ImportError:
        sys.path.insert(0, os.path.join(HERE, "ray_shim"))
        import ray  # noqa: F401

        substrate = "ray-shim"
    sys.path.insert(0, ref_dir)
    import fed  # the reference package

    if not os.path.abspath(fed.__file__).startswith(os.path.abspath(ref_dir)):
        raise ImportError(f"`import fed` resolved to {fed.__file__}, not to baseline/_ref")
    return fed, ray, substrate


# ---- a plain PyTorch ResNet-50 (He et al. 2015, torchvision layout; written here so that the
# reference arm shares no code with the product) ------------------------------------------------
def resnet50(num_classes=NUM_CLASSES):
    import torch.nn as nn

    class Bottleneck(nn.Module):
        def __init__(self, cin, width, stride):
            super().__init__()
            cout = width * 4
            self.c1 = nn.Conv2d(cin, width, 1, bias=False)
            self.b1 = nn.BatchNorm2d(width)
            self.c2 = nn.Conv2d(width, width, 3, stride, 1, bias=False)
            self.b2 = nn.BatchNorm2d(width)
            self.c3 = nn.Conv2d(width, cout, 1, bias=False)
            self.b3 = nn.BatchNorm2d(cout)
            self.act = nn.ReLU(inplace=True)
            self.down = None
            if stride != 1 or cin != cout:
                self.down = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

        def forward(self, x):
            idt = x if self.down is None else self.down(x)
            y = self.act(self.b1(self.c1(x)))
            y = self.act(self.b2(self.c2(y)))
            y = self.b3(self.c3(y))
            return self.act(y + idt)

    layers = [nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True), nn.MaxPool2d(3, 2, 1)]
    cin = 64
    for width, blocks, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
        for b in range(blocks):
            layers.append(Bottleneck(cin, width, stride if b == 0 else 1))
            cin = width * 4
    layers += [nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(cin, num_classes)]
    return nn.Sequential(*layers)


class RefWorker:
    """One party's trainer, as a user of the reference would write it: PyTorch model on the
    party's GPU, ``torch.optim.SGD``, weights exchanged as one flat CPU tensor."""

    def __init__(self, rank, device_index, seed=0):
        import torch

        self.torch = torch
        self.dev = torch.device(f"cuda:{device_index}")
        torch.cuda.set_device(self.dev)
        torch.manual_seed(seed)  # identical initial weights in every party
        self.model = resnet50().to(self.dev).to(memory_format=torch.channels_last).to(torch.bfloat16)
        self.opt = torch.optim.SGD(self.model.parameters(), lr=0.05, momentum=0.9, weight_decay=5e-5)
        g = torch.Generator().manual_seed(1000 + rank)
        nbuf = 8
        self._x = [torch.randn(BATCH, 3, IMG, IMG, generator=g).to(torch.bfloat16).pin_memory() for _ in range(nbuf)]
        self._y = [torch.randint(0, NUM_CLASSES, (BATCH,), generator=g).pin_memory() for _ in range(nbuf)]
        self._i = 0
        self._ev = {}
        self.h2d_bytes_per_step = self._x[0].numel() * 2 + self._y[0].numel() * 8
        self.numel = sum(p.numel() for p in self.model.parameters())

    def local_round(self, steps):
        torch = self.torch
        loss = None
        for _ in range(steps):
            j = self._i % len(self._x)
            self._i += 1
            x = self._x[j].to(self.dev, non_blocking=True).contiguous(memory_format=torch.channels_last)
            y = self._y[j].to(self.dev, non_blocking=True)
            loss = torch.nn.functional.cross_entropy(self.model(x).float(), y)
            self.opt.zero_grad(set_to_none=True)
            loss.backward()
            self.opt.step()
        return float(loss)  # D2H of the round's result

    def get_weights(self):
        torch = self.torch
        with torch.no_grad():
            return torch.cat([p.detach().reshape(-1) for p in self.model.parameters()]).cpu()

    def set_weights(self, flat):
        torch = self.torch
        with torch.no_grad():
            flat = flat.to(self.dev, non_blocking=True)
            off = 0
            for p in self.model.parameters():
                n = p.numel()
                p.copy_(flat[off:off + n].view_as(p))
                off += n
        return True

    def mark(self, name):
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(self.torch.cuda.current_stream())
        self._ev[name] = ev
        return True

    def elapsed_ms(self, a, b):
        self._ev[b].synchronize()
        return self._ev[a].elapsed_time(self._ev[b])

    def sync(self):
        self.torch.cuda.synchronize()
        return True

    def checksum(self):
        return float(self.get_weights().double().sum())


def mean_weights(*ws):
    import torch

    return torch.stack([w.float() for w in ws]).mean(0).to(ws[0].dtype)


def many_tiny_tasks(fed, parties, iters):
    @fed.remote
    class Tiny:
        def f(self):
            return None

        def agg(self, *xs):
            return None

    actors = {p: Tiny.party(p).remote() for p in parties}
    for _ in range(20):
        fed.get(actors[parties[0]].agg.remote(*[actors[p].f.remote() for p in parties]))
    t0 = time.perf_counter()
    for _ in range(iters):
        fed.get(actors[parties[0]].agg.remote(*[actors[p].f.remote() for p in parties]))
    dt = time.perf_counter() - t0
    n = len(parties)
    return {"iterations": iters, "per_iteration_ms": dt / iters * 1e3, "tasks_per_sec": (n + 1) * iters / dt,
            "tasks_per_iteration": n + 1, "cross_party_messages_per_iteration": 2 * (n - 1)}


def run(args, clock_sampler_cls=None):
    rank = _env_int("RANK", 0)
    world = _env_int("WORLD_SIZE", 1)
    local_rank = _env_int("LOCAL_RANK", 0)
    n = args.gpus
    try:
        fed, ray, substrate = import_reference()
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"{type(e).__name__}: {e}"[:400]}), flush=True)
        return
    assert world == n, f"--gpus {n} but WORLD_SIZE={world}; launch with torchrun --nproc-per-node {n}"
    import torch

    if not torch.cuda.is_available():
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "no CUDA device on this machine (the reference "
                              "itself imports fine)", "reference_importable": True, "substrate": substrate}),
                  flush=True)
        return
    torch.cuda.set_device(local_rank)
    # A hang of the reference (or of the stand-in under it) must not cost the line: past the deadline
    # every rank leaves; rank 0 prints the device-timed headline if it was already measured.
    import threading

    deadline_s = float(os.environ.get("RAYFED_B200_REF_ARM_DEADLINE_S", "900"))
    partial = {}

    def bail():
        if rank == 0:
            if "value" in partial:
                print(json.dumps(dict(partial, watchdog=f"the reference arm exceeded {deadline_s:.0f} s after its "
                                      "device-timed region: partial result")), flush=True)
            else:
                print(json.dumps({"impl": "reference", "unavailable": f"the reference arm did not finish its timed "
                                  f"region within {deadline_s:.0f} s"}), flush=True)
        os._exit(0)

    watchdog = threading.Timer(deadline_s, bail)
    watchdog.daemon = True
    watchdog.start()
    base = _env_int("MASTER_PORT", 29500) + 300
    parties = [f"p{i}" for i in range(n)]
    me = parties[rank]
    addresses = {f"p{i}": f"127.0.0.1:{base + i}" for i in range(n)}
    ray.init(address="local", include_dashboard=False) if substrate == "ray" else ray.init()
    fed.init(addresses=addresses, party=me, logging_level="warning",
             config={"cross_silo_comm": {"messages_max_size_in_bytes": 1 << 30, "timeout_in_ms": 120000,
                                         "exit_on_sending_failure": True},
                     "barrier_on_initializing": True},
             job_name=f"bench-ref-{base}")

    Worker = fed.remote(RefWorker)
    workers = {p: Worker.party(p).remote(i, i % torch.cuda.device_count()) for i, p in enumerate(parties)}
    mean = fed.remote(mean_weights)
    L, K_steps, W = args.local_steps, args.steps, max(args.warmup, 3)

    def submit_round():
        losses = [workers[p].local_round.remote(L) for p in parties]
        ws = [workers[p].get_weights.remote() for p in parties]
        m = mean.party(parties[0]).remote(*ws)
        for p in parties:
            workers[p].set_weights.remote(m)
        return losses

    for _ in range(W):
        fed.get(submit_round())
    fed.get([workers[p].sync.remote() for p in parties])
    sampler = clock_sampler_cls(local_rank) if (clock_sampler_cls and rank == 0) else None
    if sampler:
        sampler.start()

    # device-timed region: CUDA events on each party's stream around K rounds, max over parties
    for p in parties:
        workers[p].mark.remote("t0")
    for _ in range(K_steps):
        submit_round()
    for p in parties:
        workers[p].mark.remote("t1")
    ms = fed.get([workers[p].elapsed_ms.remote("t0", "t1") for p in parties])
    dev_ms = max(ms) / K_steps
    partial.update({"metric": "fedavg_resnet50_bf16_samples_per_sec", "value": n * L * BATCH / (dev_ms * 1e-3),
                    "unit": "samples/s", "n_gpus": n, "steps": K_steps, "warmup": W, "ms_per_step": dev_ms,
                    "higher_is_better": True, "scaling": "weak", "dtype": "bf16", "impl": "reference",
                    "substrate": substrate})

    # end to end: wall clock around K rounds incl. the per-round fed.get of the losses
    fed.get([workers[p].sync.remote() for p in parties])
    t0 = time.perf_counter()
    last = None
    for _ in range(K_steps):
        last = fed.get(submit_round())
    fed.get([workers[p].sync.remote() for p in parties])
    e2e_s = time.perf_counter() - t0
    ident = fed.remote(lambda x: x)
    e2e_s = max(fed.get([ident.party(p).remote(e2e_s) for p in parties])) if n > 1 else e2e_s
    clocks = sampler.stop() if sampler else None
    sums = fed.get([workers[p].checksum.remote() for p in parties])
    tiny = many_tiny_tasks(fed, parties, args.tiny_iters)

    samples = n * L * BATCH
    out = {
        "metric": "fedavg_resnet50_bf16_samples_per_sec", "value": samples / (dev_ms * 1e-3), "unit": "samples/s",
        "n_gpus": n, "steps": K_steps, "warmup": W, "ms_per_step": dev_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic (random 224x224 images in pinned host memory, random-init ResNet-50)",
        "impl": "reference", "substrate": substrate,
        "substrate_note": ("unmodified reference code over its own gRPC proxies; Ray replaced by baseline/ray_shim "
                           "(no Ray wheel offline): a lower bound on the reference's cost" if substrate != "ray"
                           else "unmodified reference on Ray"),
        "rounds_per_sec": 1e3 / dev_ms,
        "config": {"model": "ResNet-50 (25.56M params)", "global_batch": n * BATCH, "seq_len": None, "image": IMG,
                   "local_steps_per_round": L, "per_party_batch": BATCH,
                   "parallelism": f"fedavg-dp{n} (one party per GPU)",
                   "step_definition": "1 FedAvg round = local_steps x (H2D batch + fwd/bwd + torch.optim.SGD) + "
                                      "reference-style aggregation (weights -> p0 over gRPC, mean, -> all)",
                   "l2_policy": "inputs larger than L2 (8 distinct pinned batches rotate; >1 GB activations/step)",
                   "transport": "grpc (reference GrpcSenderProxy/GrpcReceiverProxy)"},
        "e2e": {"value": samples * K_steps / e2e_s, "unit": "samples/s", "rounds_per_sec": K_steps / e2e_s,
                "h2d_bytes_per_step": (BATCH * 3 * IMG * IMG * 2 + BATCH * 8) * L, "d2h_bytes_per_step": 4 * n,
                "last_losses": last},
        "agg_check": {"parties_equal": len({round(s, 3) for s in sums}) == 1, "checksums": sums},
        "clocks": clocks, "many_tiny_tasks": tiny,
    }
    fed.shutdown()
    ray.shutdown()
    watchdog.cancel()
    if rank == 0:
        print(json.dumps(out), flush=True)
