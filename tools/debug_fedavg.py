"""2-party debugging harness (GPU): isolates transport vs aggregate vs bucket backend."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def run(rank, world, base_port):
    import rayfed_b200 as fed
    from rayfed_b200.fl import FedAggregator, FlatBucket, device_context
    from rayfed_b200.models import MLP

    parties = [f"p{i}" for i in range(world)]
    me = parties[rank]
    torch.cuda.set_device(rank)
    fed.init(addresses={p: f"127.0.0.1:{base_port + i}" for i, p in enumerate(parties)}, party=me,
             logging_level="warning",
             config={"device": f"cuda:{rank}", "barrier_on_initializing": True,
                     "cross_silo_comm": {"transport": "nvlink", "retry_initial_backoff_s": 0.05,
                                         "retry_max_attempts": 100, "exit_on_sending_failure": True}})
    out = {"rank": rank}

    # ---- 1. transport: sequences of tensors both ways, checksums compared on both sides -------
    @fed.remote
    class Box:
        def __init__(self, seed):
            self.g = torch.Generator(device="cuda").manual_seed(seed)
            self.keep = []

        def make(self, n):
            t = torch.randn(n, device="cuda", generator=self.g).to(torch.bfloat16)
            self.keep.append(t)
            return t

        def clone_of(self, i):
            return self.keep[i].clone()

        def checksum(self, i):
            return float(self.keep[i].float().sum())

    boxes = {p: Box.party(p).remote(100 + i) for i, p in enumerate(parties)}
    sizes = [700_000, 1_400_000, 32, 5_000_000]
    objs = [[boxes[p].make.remote(n) for p in parties] for n in sizes]
    res = []
    for k, row in enumerate(objs):
        vals = fed.get(row)
        sums = fed.get([boxes[p].checksum.remote(k) for p in parties])
        res.append([(float(v.float().sum()), s) for v, s in zip(vals, sums)])
    out["transport_make"] = res
    clones = fed.get([boxes[p].clone_of.remote(1) for p in parties])
    sums = fed.get([boxes[p].checksum.remote(1) for p in parties])
    out["transport_clone"] = [(float(v.float().sum()), s) for v, s in zip(clones, sums)]

    # ---- 2. aggregate on ipc bucket vs vmm(+mc) bucket ------------------------------------------
    @fed.remote
    class Agg:
        def __init__(self, rank, multicast, tag):
            group, heap, channel = device_context()
            torch.manual_seed(0)
            self.model = MLP(256, 512, 64, depth=3).cuda().to(torch.bfloat16)
            self.bucket = FlatBucket(self.model, heap, name=f"dbg{tag}", n_parties=group.size, multicast=multicast)
            g = torch.Generator(device="cuda").manual_seed(500 + rank)
            self.bucket.flat_w.copy_(torch.randn(self.bucket.numel, device="cuda", generator=g).to(torch.bfloat16))
            self.agg = FedAggregator(self.bucket, group, channel, kind="fedavg", heap=heap)
            self.backend = self.bucket.buffer.backend

        def weights(self):
            return self.bucket.flat_w.clone()

        def aggregate(self):
            self.agg.aggregate()
            return True

        def info(self):
            return self.backend, self.agg.uses_multicast, self.agg.status()

    for multicast, tag in ((False, "ipc"), (True, "vmm")):
        aggs = {p: Agg.party(p).remote(i, multicast, tag) for i, p in enumerate(parties)}
        before = [t.float().cpu() for t in fed.get([aggs[p].weights.remote() for p in parties])]
        for p in parties:
            aggs[p].aggregate.remote()
        after = [t.float().cpu() for t in fed.get([aggs[p].weights.remote() for p in parties])]
        exp = (sum(before) / len(before)).to(torch.bfloat16).float()
        info = fed.get([aggs[p].info.remote() for p in parties])
        out[f"agg_{tag}"] = {
            "info": info,
            "before_sums": [float(b.sum()) for b in before],
            "after_err": [float((a - exp).abs().max()) for a in after],
            "after_sums": [float(a.sum()) for a in after], "exp_sum": float(exp.sum()),
        }
    # ---- 3. the FedAvgWorker flow of tests/test_gpu_fed_nvlink.py with sender-side checksums ------
    from rayfed_b200.fl import FedAvgWorker
    from rayfed_b200.models import resnet18

    class W(FedAvgWorker):
        def snap(self, mode):
            if mode == "device_sync":
                torch.cuda.synchronize()
            elif mode == "lane_sync":
                torch.cuda.current_stream().synchronize()
            c = self.bucket.flat_w.clone()
            return c

        def live_sum(self):
            torch.cuda.synchronize()
            return float(self.bucket.flat_w.float().sum())

    def batch_fn(r, i, _rank=rank):
        g = torch.Generator().manual_seed(100 * _rank + 10 * r + i)
        return torch.randn(8, 3, 32, 32, generator=g), torch.randint(0, 10, (8,), generator=g)

    for mode in ("none", "lane_sync", "device_sync", "none"):
        Wk = fed.remote(W)
        ws = {p: Wk.party(p).remote(lambda: resnet18(10, width=16), parties=parties, party=p, batch_fn=batch_fn,
                                    lr=0.05, kind="fedavg") for p in parties}
        losses = [ws[p].local_round.remote(2) for p in parties]
        b = fed.get([ws[p].snap.remote(mode) for p in parties])
        for p in parties:
            ws[p].aggregate.remote()
        a = fed.get([ws[p].snap.remote(mode) for p in parties])
        live = fed.get([ws[p].live_sum.remote() for p in parties])
        bt = [x.float().cpu() for x in b]
        at = [x.float().cpu() for x in a]
        exp = ((bt[0] + bt[1]) * 0.5).to(torch.bfloat16).float()
        out.setdefault("worker", []).append({
            "mode": mode,
            "before_sums": [float(t.sum()) for t in bt],
            "after_sums": [float(t.sum()) for t in at],
            "live_sums_after": live, "exp_sum": float(exp.sum()),
            "after_err_vs_exp": [float((t - exp).abs().max()) for t in at],
            "after_diff": float((at[0] - at[1]).abs().max()),
        })
    fed.shutdown()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/debug_fedavg_{rank}.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(run, args=(world, 14100), nprocs=world, join=True)
