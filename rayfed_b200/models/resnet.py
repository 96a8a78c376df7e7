"""ResNet (He et al. 2015) — the FedAvg image model of BASELINE.json's
"FedAvg ResNet-50 bf16" config.  Plain PyTorch modules (cuDNN convolutions);
what is federated-specific — flat parameter buckets in the symmetric heap, the
fused local optimizer and the fused aggregate — lives in :mod:`rayfed_b200.fl`.
ResNet-50 has 25.56 M parameters (≈51 MB in bf16)."""
from __future__ import annotations

import torch
from torch import nn


def _conv3x3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 3, stride, 1, bias=False)


def _conv1x1(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 1, stride, 0, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, cin, width, stride=1, down=None):
        super().__init__()
        self.conv1, self.bn1 = _conv3x3(cin, width, stride), nn.BatchNorm2d(width)
        self.conv2, self.bn2 = _conv3x3(width, width), nn.BatchNorm2d(width)
        self.relu = nn.ReLU(inplace=True)
        self.down = down

    def forward(self, x):
        idt = x if self.down is None else self.down(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return self.relu(y + idt)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin, width, stride=1, down=None):
        super().__init__()
        self.conv1, self.bn1 = _conv1x1(cin, width), nn.BatchNorm2d(width)
        self.conv2, self.bn2 = _conv3x3(width, width, stride), nn.BatchNorm2d(width)
        self.conv3, self.bn3 = _conv1x1(width, width * 4), nn.BatchNorm2d(width * 4)
        self.relu = nn.ReLU(inplace=True)
        self.down = down

    def forward(self, x):
        idt = x if self.down is None else self.down(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return self.relu(y + idt)


class ResNet(nn.Module):
    def __init__(self, block, depths, num_classes=1000, width=64):
        super().__init__()
        self.cin = width
        self.stem = nn.Sequential(
            nn.Conv2d(3, width, 7, 2, 3, bias=False), nn.BatchNorm2d(width), nn.ReLU(inplace=True),
            nn.MaxPool2d(3, 2, 1),
        )
        self.layer1 = self._stage(block, width, depths[0], 1)
        self.layer2 = self._stage(block, width * 2, depths[1], 2)
        self.layer3 = self._stage(block, width * 4, depths[2], 2)
        self.layer4 = self._stage(block, width * 8, depths[3], 2)
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(width * 8 * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        for m in self.modules():  # zero-init the last BN of each residual branch
            if isinstance(m, Bottleneck):
                nn.init.zeros_(m.bn3.weight)
            elif isinstance(m, BasicBlock):
                nn.init.zeros_(m.bn2.weight)

    def _stage(self, block, width, n, stride):
        down = None
        if stride != 1 or self.cin != width * block.expansion:
            down = nn.Sequential(_conv1x1(self.cin, width * block.expansion, stride),
                                 nn.BatchNorm2d(width * block.expansion))
        layers = [block(self.cin, width, stride, down)]
        self.cin = width * block.expansion
        layers += [block(self.cin, width) for _ in range(1, n)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.stem(x)
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(self.pool(x), 1))


def resnet18(num_classes=1000, width=64):
    return ResNet(BasicBlock, [2, 2, 2, 2], num_classes, width)


def resnet50(num_classes=1000, width=64):
    return ResNet(Bottleneck, [3, 4, 6, 3], num_classes, width)
