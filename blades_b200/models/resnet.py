"""ResNet-18 / ResNet-50 for the BASELINE configs (the reference ships neither --
SURVEY section 0; parameter layout follows torchvision so d = 11 181 642 for ResNet-18 /
10 classes and 23 712 932 for ResNet-50 / 100 classes, and torchvision state-dicts load).

Normalisation decision: the reference's update vector contains only
trainable parameters and BatchNorm running statistics are never aggregated (server.py:
66-75, client.py:216-228), so with stock BatchNorm the global model's running stats
would stay at their initial values forever.  Default here is therefore ``norm='batch'``
with ``track_running_stats=False``: per-client batch statistics in training ("ghost"
BN over each client's batch) and batch statistics at eval; no buffers exist, the
trainable-parameter vector is identical to torchvision's.  ``norm='group'`` (GroupNorm-32,
same parameter count) and ``norm='batch_running'`` (stock BN) are available.
"""
from __future__ import annotations

from typing import Callable, List, Type

import torch
import torch.nn as nn

__all__ = ["ResNet", "resnet18", "resnet50", "BasicBlock", "Bottleneck"]


def _norm_factory(kind: str) -> Callable[[int], nn.Module]:
    if kind == "batch":
        return lambda c: nn.BatchNorm2d(c, track_running_stats=False)
    if kind == "batch_running":
        return lambda c: nn.BatchNorm2d(c)
    if kind == "group":
        return lambda c: nn.GroupNorm(min(32, c), c)
    raise ValueError(f"unknown norm {kind!r}")


def _conv3x3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 3, stride, 1, bias=False)


def _conv1x1(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 1, stride, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, norm=None):
        super().__init__()
        self.conv1 = _conv3x3(inplanes, planes, stride)
        self.bn1 = norm(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _conv3x3(planes, planes)
        self.bn2 = norm(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + idt)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, norm=None):
        super().__init__()
        self.conv1 = _conv1x1(inplanes, planes)
        self.bn1 = norm(planes)
        self.conv2 = _conv3x3(planes, planes, stride)
        self.bn2 = norm(planes)
        self.conv3 = _conv1x1(planes, planes * self.expansion)
        self.bn3 = norm(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return self.relu(out + idt)


class ResNet(nn.Module):
    def __init__(self, block: Type[nn.Module], layers: List[int], num_classes: int = 10,
                 norm: str = "batch", in_channels: int = 3, zero_init_residual: bool = False):
        super().__init__()
        self._norm = _norm_factory(norm)
        self.norm_kind = norm
        self.inplanes = 64
        self.conv1 = nn.Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = self._norm(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        self.zero_init_residual = zero_init_residual
        self.reset_parameters_()

    def reset_parameters_(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if self.zero_init_residual:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    nn.init.constant_(m.bn3.weight, 0)
                elif isinstance(m, BasicBlock):
                    nn.init.constant_(m.bn2.weight, 0)

    def _make_layer(self, block, planes, blocks, stride=1):
        down = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            down = nn.Sequential(_conv1x1(self.inplanes, planes * block.expansion, stride),
                                 self._norm(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, down, self._norm)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes, norm=self._norm) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(self.avgpool(x), 1))


def resnet18(num_classes: int = 10, norm: str = "batch", **kw) -> ResNet:
    return ResNet(BasicBlock, [2, 2, 2, 2], num_classes=num_classes, norm=norm, **kw)


def resnet50(num_classes: int = 100, norm: str = "batch", **kw) -> ResNet:
    return ResNet(Bottleneck, [3, 4, 6, 3], num_classes=num_classes, norm=norm, **kw)
