"""Small dense models for tests and the split-learning example."""
from torch import nn


class MLP(nn.Module):
    def __init__(self, d_in=32, d_hidden=64, d_out=10, depth=2):
        super().__init__()
        layers, d = [], d_in
        for _ in range(depth):
            layers += [nn.Linear(d, d_hidden), nn.ReLU()]
            d = d_hidden
        layers.append(nn.Linear(d, d_out))
        self.net = nn.Sequential(*layers)

    def forward(self, x):
        return self.net(x)


class SplitBottom(nn.Module):
    """Feature extractor held by the data-owning party in split learning."""

    def __init__(self, d_in=32, d_cut=16):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(d_in, 64), nn.ReLU(), nn.Linear(64, d_cut), nn.ReLU())

    def forward(self, x):
        return self.net(x)


class SplitTop(nn.Module):
    """Head held by the label-owning party."""

    def __init__(self, d_cut=16, d_out=10):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(d_cut, 32), nn.ReLU(), nn.Linear(32, d_out))

    def forward(self, h):
        return self.net(h)
