"""FedAvg of a small ResNet over the NVLink transport — one party per GPU.

    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 2 examples/fedavg_resnet.py
"""
import os

import torch

import rayfed_b200 as fed
from rayfed_b200.fl import FedAvgWorker
from rayfed_b200.models import resnet18


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    parties = [f"p{i}" for i in range(world)]
    addresses = {p: f"127.0.0.1:{12000 + i}" for i, p in enumerate(parties)}
    torch.cuda.set_device(local)
    fed.init(addresses=addresses, party=parties[rank], logging_level="warning",
             config={"device": f"cuda:{local}", "cross_silo_comm": {"transport": "nvlink"},
                     "barrier_on_initializing": True})

    def batch_fn(r, i, _rank=rank):
        g = torch.Generator().manual_seed(1000 * _rank + 10 * r + i)
        return torch.randn(64, 3, 32, 32, generator=g), torch.randint(0, 10, (64,), generator=g)

    Worker = fed.remote(FedAvgWorker)
    workers = {p: Worker.party(p).remote(lambda: resnet18(10, width=32), parties=parties, party=p,
                                         batch_fn=batch_fn, lr=0.05) for p in parties}
    for rnd in range(5):
        losses = [workers[p].local_round.remote(4) for p in parties]      # local SGD on each party's data
        for p in parties:
            workers[p].aggregate.remote()                                 # ONE fused kernel per party
        print(parties[rank], "round", rnd, [round(float(x), 3) for x in fed.get(losses)])
    fed.shutdown()


if __name__ == "__main__":
    main()
