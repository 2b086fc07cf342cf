"""Dilated ResNet encoders (single- and two-stream) on the B200 kernel library.

Host-side mirror of the reference's `models/extractors.py` and
`models/confusion_maximization/models/extractors.py` (class names, constructor signatures, parameter /
buffer names and shapes are identical, so reference checkpoints load unchanged).  The nn.Conv2d /
nn.BatchNorm2d children are parameter holders only -- their own forward is never called; every module
runs through `engine` on NHWC views.  Reference: cm/models/extractors.py:66-198, 390-394.
"""
import math

import torch
import torch.nn as nn
from torch.utils import model_zoo

from . import engine as E
from .engine import ACT_NONE, ACT_RELU, Act

model_urls = {
    'resnet18': 'https://download.pytorch.org/models/resnet18-5c106cde.pth',
    'resnet34': 'https://download.pytorch.org/models/resnet34-333f7ec4.pth',
    'resnet50': 'https://download.pytorch.org/models/resnet50-19c8e357.pth',
    'resnet101': 'https://download.pytorch.org/models/resnet101-5d3b4d8f.pth',
    'resnet152': 'https://download.pytorch.org/models/resnet152-b121ed2d.pth',
}


def load_weights_sequential(target, source_state):
    model_to_load = {k: v for k, v in source_state.items() if k in target.state_dict().keys()}
    target.load_state_dict(model_to_load)


def conv3x3(in_planes, out_planes, stride=1, dilation=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=dilation, dilation=dilation, bias=False)


class _KernelModule(nn.Module):
    """Common tensor-facing forward: NCHW tensor -> NHWC Act -> kernels -> NCHW view."""
    precision = None      # None -> inherit engine.DEFAULT_PRECISION

    def _dtype(self):
        return E.precision_dtype(self.precision or E.DEFAULT_PRECISION)

    def forward(self, x):
        return self._run(E.from_nchw(x, self._dtype())).nchw()


class BasicBlock(_KernelModule):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1):
        super(BasicBlock, self).__init__()
        self.conv1 = conv3x3(inplanes, planes, stride=stride, dilation=dilation)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes, stride=1, dilation=dilation)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def _run(self, x: Act, out: Act = None) -> Act:
        o = E.conv_bn_act(x, self.conv1, self.bn1, ACT_RELU)
        res = x if self.downsample is None else E.conv_bn_act(x, self.downsample[0], self.downsample[1], ACT_NONE)
        return E.conv_bn_act(o, self.conv2, self.bn2, ACT_RELU, residual=res, out=out)


class Bottleneck(_KernelModule):
    """cm/models/extractors.py:66-102: 1x1 -> BN -> ReLU -> 3x3 -> BN -> ReLU -> 1x1 -> BN -> (+res) -> ReLU.
    In eval mode that is 3 (4 with a downsample branch) launches: BN, residual add and ReLU live in the conv
    epilogues."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1):
        super(Bottleneck, self).__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, dilation=dilation, padding=dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def _run(self, x: Act, out: Act = None) -> Act:
        o = E.conv_bn_act(x, self.conv1, self.bn1, ACT_RELU)
        o = E.conv_bn_act(o, self.conv2, self.bn2, ACT_RELU)
        res = x if self.downsample is None else E.conv_bn_act(x, self.downsample[0], self.downsample[1], ACT_NONE)
        return E.conv_bn_act(o, self.conv3, self.bn3, ACT_RELU, residual=res, out=out)


class ResNet(_KernelModule):
    """cm/models/extractors.py:105-198.  `late_fusion=True` adds the IR stream (conv1_2, bn1_2, layer1_2,
    layer2_2); the two streams write into channel halves of shared NHWC buffers, so the reference's three
    torch.cat calls cost nothing.  With `late_fusion=False, in_channels=3` this is the top-level
    models/extractors.py:105-154 ResNet (identical state_dict)."""

    def __init__(self, block, layers=(3, 4, 23, 3), late_fusion=False, in_channels=3):
        self.inplanes = 64
        self.late_fusion = late_fusion
        super(ResNet, self).__init__()
        if not late_fusion:
            self.conv1 = nn.Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
            self.bn1 = nn.BatchNorm2d(64)
            self.relu = nn.ReLU(inplace=True)
        else:
            self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)       # RGB
            self.bn1 = nn.BatchNorm2d(64)
            self.relu = nn.ReLU(inplace=True)
            self.conv1_2 = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)     # IR
            self.bn1_2 = nn.BatchNorm2d(64)
            self.relu_2 = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)

        channel_copy = self.inplanes
        self.layer1 = self._make_layer(block, 64, layers[0])
        if self.late_fusion:
            self.inplanes = channel_copy
            self.layer1_2 = self._make_layer(block, 64, layers[0])
        channel_copy = self.inplanes
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        if self.late_fusion:
            self.inplanes = channel_copy
            self.layer2_2 = self._make_layer(block, 128, layers[1], stride=2)
        if self.late_fusion:
            self.inplanes = int(self.inplanes * 2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=1, dilation=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=1, dilation=4)

        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion),
            )
        layers = [block(self.inplanes, planes, stride, downsample)]      # quirk kept: block 0 gets dilation 1
        self.inplanes = planes * block.expansion
        for i in range(1, blocks):
            layers.append(block(self.inplanes, planes, dilation=dilation))
        return nn.Sequential(*layers)

    # ---- kernels
    @staticmethod
    def _layer(layer, x: Act, out: Act = None) -> Act:
        n = len(layer)
        for i, blk in enumerate(layer):
            x = blk._run(x, out if i == n - 1 else None)
        return x

    def _stem(self, x: Act, conv, bn, out: Act) -> Act:
        y = E.conv_bn_act(x, conv, bn, ACT_RELU)
        return E.maxpool3x3s2(y, out)

    def _run_taps(self, m1: Act, m2: Act = None, x5_out: Act = None):
        """-> [x5, x4, x3, x2, x1] as Acts (cm/models/extractors.py:172-198)."""
        dt, dev = m1.dtype, m1.buf.device
        two = self.late_fusion and m2 is not None
        assert self.late_fusion or m2 is None, "early fusion: fuse the inputs first (engine.fuse_inputs)"
        mult = 2 if two else 1
        h2, w2 = E.conv_out_hw(m1.h, m1.w, self.conv1)
        h4, w4 = (h2 - 1) // 2 + 1, (w2 - 1) // 2 + 1
        c1 = 64
        x1 = E.new_act(m1.n, h4, w4, c1 * mult, dt, dev)
        self._stem(m1, self.conv1, self.bn1, x1.slice(0, c1))
        if two:
            self._stem(m2, self.conv1_2, self.bn1_2, x1.slice(c1, c1))
        c2 = self.layer1[-1].bn3.num_features if hasattr(self.layer1[-1], "bn3") else self.layer1[-1].bn2.num_features
        x2 = E.new_act(m1.n, h4, w4, c2 * mult, dt, dev)
        self._layer(self.layer1, x1.slice(0, c1), x2.slice(0, c2))
        if two:
            self._layer(self.layer1_2, x1.slice(c1, c1), x2.slice(c2, c2))
        st = self.layer2[0].stride
        h8, w8 = (h4 - 1) // st + 1, (w4 - 1) // st + 1
        c3 = self.layer2[-1].bn3.num_features if hasattr(self.layer2[-1], "bn3") else self.layer2[-1].bn2.num_features
        x3 = E.new_act(m1.n, h8, w8, c3 * mult, dt, dev)
        self._layer(self.layer2, x2.slice(0, c2), x3.slice(0, c3))
        if two:
            self._layer(self.layer2_2, x2.slice(c2, c2), x3.slice(c3, c3))
        x4 = self._layer(self.layer3, x3)
        x5 = self._layer(self.layer4, x4, x5_out)
        return [x5, x4, x3, x2, x1]

    def _inputs(self, modal_1, modal_2):
        dt = self._dtype()
        if self.late_fusion:
            return E.from_nchw(modal_1, dt), (E.from_nchw(modal_2, dt) if modal_2 is not None else None)
        return E.fuse_inputs(modal_1, modal_2, dt), None

    def forward(self, modal_1, modal_2=None):
        taps = self._run_taps(*self._inputs(modal_1, modal_2))
        return [t.nchw() for t in taps]


def squeezenet(pretrained=True):
    raise NotImplementedError("squeezenet backend is outside the B200 hot path (SURVEY.md section 2, row 1)")


def densenet(pretrained=True):
    raise NotImplementedError("densenet backend is outside the B200 hot path (SURVEY.md section 2, row 1)")


def resnet18(pretrained=True, late_fusion=False, in_channels=3):
    model = ResNet(BasicBlock, [2, 2, 2, 2], late_fusion, in_channels)
    if pretrained:
        load_weights_sequential(model, model_zoo.load_url(model_urls['resnet18']))
    return model


def resnet34(pretrained=True, late_fusion=False, in_channels=3):
    model = ResNet(BasicBlock, [3, 4, 6, 3], late_fusion, in_channels)
    if pretrained:
        load_weights_sequential(model, model_zoo.load_url(model_urls['resnet34']))
    return model


def resnet50(pretrained=True, late_fusion=False, in_channels=3):
    model = ResNet(Bottleneck, [3, 4, 6, 3], late_fusion, in_channels)
    if pretrained:
        load_weights_sequential(model, model_zoo.load_url(model_urls['resnet50']))
    return model


def resnet101(pretrained=True, late_fusion=False, in_channels=3):
    model = ResNet(Bottleneck, [3, 4, 23, 3], late_fusion, in_channels)
    if pretrained:
        load_weights_sequential(model, model_zoo.load_url(model_urls['resnet101']))
    return model


def resnet152(pretrained=True, late_fusion=False, in_channels=3):
    model = ResNet(Bottleneck, [3, 8, 36, 3], late_fusion, in_channels)
    if pretrained:
        load_weights_sequential(model, model_zoo.load_url(model_urls['resnet152']))
    return model
