"""Split learning across two parties (features at alice, labels at bob).

    python examples/split_learning.py alice &  python examples/split_learning.py bob
"""
import sys

import torch

import rayfed_b200 as fed
from rayfed_b200.models import SplitBottom, SplitTop
from rayfed_b200.parallel.split import SplitBottomWorker, SplitTopWorker


@fed.remote
def make_features(step):
    g = torch.Generator().manual_seed(step)
    return torch.randn(64, 32, generator=g)


@fed.remote
def make_labels(step):
    g = torch.Generator().manual_seed(step)
    x = torch.randn(64, 32, generator=g)
    return (x[:, :10].argmax(1)).long()  # a learnable function of the (hidden) features


def main(party):
    fed.init(addresses={"alice": "127.0.0.1:11112", "bob": "127.0.0.1:11111"}, party=party,
             logging_level="warning")
    bottom = fed.remote(SplitBottomWorker).party("alice").remote(lambda: SplitBottom(32, 16))
    top = fed.remote(SplitTopWorker).party("bob").remote(lambda: SplitTop(16, 10))
    for step in range(30):
        h = bottom.forward.remote(make_features.party("alice").remote(step))   # alice -> bob: activations
        loss, grad_h = top.step.options(num_returns=2).remote(h, make_labels.party("bob").remote(step))
        bottom.backward.remote(grad_h)                                         # bob -> alice: gradients
        if step % 10 == 0:
            print(party, step, fed.get(loss))
    fed.shutdown()


if __name__ == "__main__":
    main(sys.argv[1])
