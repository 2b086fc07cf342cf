"""Network factory: drop-in for the reference's two `build_net.build_network` functions.

  models/build_net.py:18-28                          build_network(snapshot, backend) -> (net, epoch)
  models/confusion_maximization/models/build_net.py  build_network(snapshot, backend, in_channels, late_fusion) -> net

Same snapshot convention (`<name>_<epoch>` file holding a state_dict) and the same unconditional `.cuda()`.
"""
import os

import torch

from .pspnet import PSPNet, PSPNetRGB

models = {
    'squeezenet': lambda: PSPNetRGB(sizes=(1, 2, 3, 6), psp_size=512, deep_features_size=256, backend='squeezenet'),
    'densenet': lambda: PSPNetRGB(sizes=(1, 2, 3, 6), psp_size=1024, deep_features_size=512, backend='densenet'),
    'resnet18': lambda: PSPNetRGB(sizes=(1, 2, 3, 6), psp_size=512, deep_features_size=256, backend='resnet18'),
    'resnet34': lambda: PSPNetRGB(sizes=(1, 2, 3, 6), psp_size=512, deep_features_size=256, backend='resnet34'),
    'resnet50': lambda: PSPNetRGB(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50'),
    'resnet101': lambda: PSPNetRGB(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet101'),
    'resnet152': lambda: PSPNetRGB(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet152')
}

_UNSET = object()


def build_network(snapshot, backend, in_channels=_UNSET, late_fusion=_UNSET):
    epoch = 0
    backend = backend.lower()
    heatnet_signature = in_channels is not _UNSET or late_fusion is not _UNSET
    if heatnet_signature:
        # cm/models/build_net.py:21 always builds ResNet-50, pretrained=False, whatever `backend` says
        net = PSPNet(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50',
                     in_channels=3 if in_channels is _UNSET else in_channels, pretrained=False,
                     late_fusion=False if late_fusion is _UNSET else late_fusion)
    else:
        net = models[backend]()          # KeyError for an unknown backend, like the reference
    if snapshot is not None:
        _, epoch = os.path.basename(snapshot).split('_')
        epoch = int(epoch)
        net.load_state_dict(torch.load(snapshot))
    net = net.cuda()
    return net if heatnet_signature else (net, epoch)
