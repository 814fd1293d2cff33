"""Host mirror of ldm/modules/id_embedding/iresnet.py:26-181,232-235: parameter containers for iresnet100
(the arithmetic is celebbasis_b200.iresnet_engine.IResNetEngine)."""
import torch
from torch import nn


class IBasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(inplanes, eps=1e-05)
        self.conv1 = nn.Conv2d(inplanes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes, eps=1e-05)
        self.prelu = nn.PReLU(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes, eps=1e-05)
        self.downsample = downsample
        self.stride = stride


class IResNet(nn.Module):
    def __init__(self, block, layers, dropout=0, num_features=512, fp16=False):
        super().__init__()
        self.fp16 = fp16
        self.inplanes = 64
        self.layer_counts = tuple(layers)
        self.conv1 = nn.Conv2d(3, 64, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(64, eps=1e-05)
        self.prelu = nn.PReLU(64)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1])
        self.layer3 = self._make_layer(block, 256, layers[2])
        self.layer4 = self._make_layer(block, 512, layers[3])
        self.bn2 = nn.BatchNorm2d(512, eps=1e-05)
        self.dropout = nn.Dropout(p=dropout, inplace=True)
        self.fc = nn.Linear(512 * 49, num_features)
        self.features = nn.BatchNorm1d(num_features, eps=1e-05)
        nn.init.constant_(self.features.weight, 1.0)
        self.features.weight.requires_grad = False

    def _make_layer(self, block, planes, blocks):
        ds = nn.Sequential(nn.Conv2d(self.inplanes, planes, 1, stride=2, bias=False), nn.BatchNorm2d(planes, eps=1e-05))
        layers = [block(self.inplanes, planes, 2, ds)]
        self.inplanes = planes
        layers += [block(planes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)


def iresnet100(pretrained=False, progress=True, **kwargs):
    return IResNet(IBasicBlock, [3, 13, 30, 3], **kwargs)
