"""Backbones of the coarse / refiner networks.  Per BASELINE.json they stay ordinary torch modules (cuDNN convs,
run in bf16); only their construction lives here.  Parameter names match the reference's modules so reference
checkpoints load with load_state_dict:
  vanilla ResNet-34 with n_input_channels   megapose/models/torchvision_resnet.py:191-374 (= torchvision's ResNet
                                            with a wider first conv), built by pose_models_cfg.py:106-113
  pre-activation WideResNet-18/34           megapose/models/wide_resnet.py:68-154
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
import torchvision
from torch import nn


def resnet34(num_classes: int = 512, n_input_channels: int = 3) -> nn.Module:
    """torchvision ResNet-34 (BasicBlock [3,4,6,3]) + fc(512 -> num_classes); conv1 takes n_input_channels."""
    net = torchvision.models.resnet.ResNet(torchvision.models.resnet.BasicBlock, [3, 4, 6, 3], num_classes=num_classes)
    if n_input_channels != 3:
        net.conv1 = nn.Conv2d(n_input_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        nn.init.kaiming_normal_(net.conv1.weight, mode="fan_out", nonlinearity="relu")
    net.n_features = num_classes
    net.n_inputs = n_input_channels
    return net


class PreActBlock(nn.Module):
    """ResNet-v2 basic block (BN-ReLU-conv twice); the shortcut conv sees the activated input."""

    expansion = 1

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: nn.Module = None):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        a = F.relu(self.bn1(x), inplace=True)
        shortcut = x if self.downsample is None else self.downsample(a)
        y = self.conv2(F.relu(self.bn2(self.conv1(a)), inplace=True))
        return y + shortcut


class WideResNet(nn.Module):
    def __init__(self, layers, width: float = 1.0, num_inputs: int = 3, maxpool: bool = True):
        super().__init__()
        chans = [int(c * width) for c in (64, 128, 256, 512)]
        self.inplanes = chans[0]
        self.conv1 = nn.Conv2d(num_inputs, chans[0], kernel_size=5, stride=2, padding=2, bias=False)
        self.bn1 = nn.BatchNorm2d(chans[0])
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1) if maxpool else nn.Identity()
        self.layer1 = self._stage(chans[0], layers[0], 1)
        self.layer2 = self._stage(chans[1], layers[1], 2)
        self.layer3 = self._stage(chans[2], layers[2], 2)
        self.layer4 = self._stage(chans[3], layers[3], 2)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        self.n_features = chans[3]
        self.n_inputs = num_inputs

    def _stage(self, planes: int, n_blocks: int, stride: int) -> nn.Sequential:
        down = None
        if stride != 1 or self.inplanes != planes:
            down = nn.Conv2d(self.inplanes, planes, kernel_size=1, stride=stride, bias=False)
        blocks = [PreActBlock(self.inplanes, planes, stride, down)]
        self.inplanes = planes
        blocks += [PreActBlock(planes, planes) for _ in range(n_blocks - 1)]
        return nn.Sequential(*blocks)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))


def WideResNet18(n_inputs: int = 3, width: float = 1.0) -> WideResNet:
    return WideResNet([2, 2, 2, 2], width, n_inputs)


def WideResNet34(n_inputs: int = 3, width: float = 1.0) -> WideResNet:
    return WideResNet([3, 4, 6, 3], width, n_inputs)


def make_backbone(backbone_str: str, n_inputs: int) -> nn.Module:
    """pose_models_cfg.py:104-122."""
    if backbone_str == "vanilla_resnet34":
        return resnet34(num_classes=512, n_input_channels=n_inputs)
    if backbone_str == "resnet34":
        return WideResNet34(n_inputs=n_inputs)
    if backbone_str == "resnet18":
        return WideResNet18(n_inputs=n_inputs)
    if "resnet34_width=" in backbone_str:
        return WideResNet34(n_inputs=n_inputs, width=int(backbone_str.split("resnet34_width=")[1]))
    raise ValueError("Unknown backbone", backbone_str)
