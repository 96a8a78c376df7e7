"""FedAvg with your own trainer actors and ``fed.aggregate`` (CPU-friendly).

    python examples/fed_aggregate.py alice &  python examples/fed_aggregate.py bob

Every party fits a linear model on its private data; each round the weights are averaged
across parties with ONE call.  With ``config={"device": "cuda:N", ...}`` the same script
uses the fused NVLink aggregate kernel instead of gather / mean / redistribute.
"""
import sys

import numpy as np

import rayfed_b200 as fed

ADDRESSES = {"alice": "127.0.0.1:11312", "bob": "127.0.0.1:11311"}
TRUE_W = np.array([2.0, -1.0, 0.5])


@fed.remote
class Trainer:
    def __init__(self, seed):
        rng = np.random.default_rng(seed)
        self.x = rng.normal(size=(256, 3))
        self.y = self.x @ TRUE_W + 0.01 * rng.normal(size=256)
        self.w = np.zeros(3)

    def local_steps(self, n, lr=0.1):
        for _ in range(n):
            grad = 2 * self.x.T @ (self.x @ self.w - self.y) / len(self.y)
            self.w = self.w - lr * grad
        return {"w": self.w, "n": float(len(self.y))}

    def set_weights(self, state):
        self.w = state["w"]
        return float(np.abs(self.w - TRUE_W).max())


def main(party, addresses=ADDRESSES, rounds=15):
    fed.init(addresses=addresses, party=party, logging_level="warning")
    parties = list(addresses)
    trainers = {p: Trainer.party(p).remote(i) for i, p in enumerate(parties)}
    err = None
    for _ in range(rounds):
        states = [trainers[p].local_steps.remote(5) for p in parties]
        averaged = fed.aggregate(states)                      # one FedObject per party, owned by that party
        errs = [trainers[p].set_weights.remote(s) for p, s in zip(parties, averaged)]
        err = max(fed.get(errs))
    print(f"[{party}] max |w - w*| after {rounds} rounds: {err:.4f}")
    fed.shutdown()
    return err


if __name__ == "__main__":
    main(sys.argv[1])
