"""Minimal repro attempt: FedAvgWorker flow with plain clone returns, varying the arena size."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def run(rank, world, base_port, arena_mb, tag):
    import rayfed_b200 as fed
    from rayfed_b200.fl import FedAvgWorker
    from rayfed_b200.models import resnet18

    parties = ["alice", "bob"]
    me = parties[rank]
    torch.cuda.set_device(rank)
    fed.init(addresses={p: f"127.0.0.1:{base_port + i}" for i, p in enumerate(parties)}, party=me,
             logging_level="warning",
             config={"device": f"cuda:{rank}", "barrier_on_initializing": True,
                     "cross_silo_comm": {"transport": "nvlink", "retry_initial_backoff_s": 0.05,
                                         "retry_max_attempts": 100, "exit_on_sending_failure": True,
                                         "device_arena_bytes": arena_mb << 20, "timeout_in_ms": 30000}})

    def batch_fn(r, i, _rank=rank):
        g = torch.Generator().manual_seed(100 * _rank + 10 * r + i)
        return torch.randn(8, 3, 32, 32, generator=g), torch.randint(0, 10, (8,), generator=g)

    out = []
    for rep in range(1):
        Worker = fed.remote(FedAvgWorker)
        ws = {p: Worker.party(p).remote(lambda: resnet18(10, width=16), parties=parties, party=p, batch_fn=batch_fn,
                                        lr=0.05, kind="fedavg", server_lr=1.0, server_momentum=0.0)
              for p in parties}
        losses = [ws[p].local_round.remote(2) for p in parties]
        before = fed.get([ws[p].get_flat_weights.remote() for p in parties])
        for p in parties:
            ws[p].aggregate.remote()
        after = fed.get([ws[p].get_flat_weights.remote() for p in parties])
        ls = fed.get(losses)
        b = [t.float().cpu() for t in before]
        a = [t.float().cpu() for t in after]
        exp = ((b[0] + b[1]) * 0.5).to(torch.bfloat16).float()
        from rayfed_b200.proxy.nvlink import nvlink_proxy

        ch = nvlink_proxy.current_channel()
        st = ch.check_status()
        out.append({"rep": rep, "status": st, "stats": dict(ch.stats), "same_after": torch.equal(a[0], a[1]),
                    "err": [float((t - exp).abs().max()) for t in a],
                    "sums_b": [float(t.sum()) for t in b], "sums_a": [float(t.sum()) for t in a]})
    fed.shutdown()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/dbg2_{tag}_{rank}.json", "w"))


if __name__ == "__main__":
    import torch.multiprocessing as mp

    arena = int(sys.argv[1])
    mp.spawn(run, args=(2, 14300 + arena + int(sys.argv[2]), arena, f"a{arena}_{sys.argv[3]}"), nprocs=2, join=True)
