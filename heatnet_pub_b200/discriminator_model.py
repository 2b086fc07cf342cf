"""Domain critics on the B200 kernel library.

Host-side mirror of the reference's `models/confusion_maximization/discriminator_model.py:35-64`
(`FCDiscriminator`: five 4x4 stride-2 convolutions with bias, LeakyReLU(0.2) between them, bilinear x32).
Bias and LeakyReLU run in the conv epilogues; the critic map is returned as FP32 NCHW like the reference's.
"""
import torch
import torch.nn as nn

from . import autograd
from . import engine as E
from .engine import ACT_LEAKY, ACT_NONE, Act
from .extractors import _KernelModule


class FCDiscriminator(_KernelModule):

    def __init__(self, num_classes, ndf=64):
        super(FCDiscriminator, self).__init__()

        self.conv1 = nn.Conv2d(num_classes, ndf, kernel_size=4, stride=2, padding=1)
        self.conv2 = nn.Conv2d(ndf, ndf*2, kernel_size=4, stride=2, padding=1)
        self.conv3 = nn.Conv2d(ndf*2, ndf*4, kernel_size=4, stride=2, padding=1)
        self.conv4 = nn.Conv2d(ndf*4, ndf*8, kernel_size=4, stride=2, padding=1)
        self.classifier = nn.Conv2d(ndf*8, 1, kernel_size=4, stride=2, padding=1)

        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        self.up_sample = nn.Upsample(scale_factor=32, mode='bilinear')

    def _run(self, x: Act) -> Act:
        slope = self.leaky_relu.negative_slope
        for conv in (self.conv1, self.conv2, self.conv3, self.conv4):
            x = E.conv_bn_act(x, conv, None, ACT_LEAKY, slope)
        ho, wo = E.conv_out_hw(x.h, x.w, self.classifier)
        if ho < 1 or wo < 1:
            raise RuntimeError(f"Calculated padded input size per channel: ({x.h + 2} x {x.w + 2}). Kernel size: (4, 4). "
                               "Kernel size can't be greater than actual input size")
        score = E.new_act(x.n, ho, wo, 1, torch.float32, x.buf.device, ld=4)      # FP32 critic map
        E.conv_bn_act(x, self.classifier, None, out=score)
        f = int(self.up_sample.scale_factor)
        return E.bilinear(score, f * score.h, f * score.w)

    def forward(self, x):
        if autograd.needs_autograd(self, x):
            return autograd.apply(self._autograd_runner, [x], self)[0]
        y = self._run(E.from_nchw(x, self._dtype()))
        return y.nchw()            # C == 1: NHWC and NCHW coincide, dense FP32 (N,1,H,W)

    def _autograd_runner(self, tape, inputs):
        xin = E.from_nchw(inputs[0], self._dtype())
        if inputs[0].requires_grad:
            tape.require(xin)
        y = self._run(xin)
        return [xin], [y], [y.nchw()]
