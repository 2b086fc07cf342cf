"""Adversarial / confusion-maximisation wrapper on the B200 kernel library.

Drop-in for the reference's `models/confusion_maximization/models/conf_segnet.py:13-140`: `create_critic`, and
`conv_segnet` with the same constructor signature, the attributes the trainer touches (`.trgb_segnet`, `.critics`,
`.phase`, `.no_conf`), `setPhase` / `setLearningModel`, the console messages and the forward dict
(`critics_a`, `critics_b`, `pred_label_a/b`, `cert_a/b`, `inter_f_b`); `state_dict()` keys are `trgb_segnet.*` and
`critics.<i>.*` like the reference's.

Only the path the north star names is built: arch='pspnet' with disc_arch='cyclegan' (FCDiscriminator).  The ResNeXt
"custom" arch, the ResNet critics, the feedback_seg DownNets, the input adapter and the certainty branch are out of
scope and raise NotImplementedError instead of silently building something else.
"""
import os

import torch.nn as nn

from . import build_net, discriminator_model, utils

# one batched seg-net / critic pass over [day; night] instead of two (PSPNet.forward_pair); HN_NO_PAIR_FORWARD=1 keeps the two calls
PAIR_FORWARD = os.environ.get("HN_NO_PAIR_FORWARD") is None

# channel counts of the six critic taps [logits, x5, x4, x3, x2, x1] (conf_segnet.py:44-50): with late fusion the three
# shallow taps are RGB || IR concatenations
_TAP_CHANNELS = {False: (13, 2048, 1024, 512, 256, 64), True: (13, 2048, 1024, 1024, 512, 128)}
_PRETRAINED_CHECKPOINT = 'models_finished/training_nc_irrgb_best.pth'      # conf_segnet.py:80
_PHASES = {"train_seg": (True, False), "train_critic": (False, True)}       # phase -> (seg net learns, critics learn)


def create_critic(disc_arch, input_num):
    """conf_segnet.py:13-20.  Only the trainer's default critic family is on the hot path."""
    if disc_arch != 'cyclegan':
        raise NotImplementedError("disc_arch=%r: only the default 'cyclegan' critics are on the B200 hot path" % (disc_arch,))
    return discriminator_model.FCDiscriminator(input_num)


def _count_input_channels(modalities):
    n = 0
    for tag, width, label in (('rgb', 3, 'RGB'), ('ir', 1, 'IR')):
        if tag in modalities:
            n += width
            print('Using %s' % label)
    print('Total numbers of input channels: %d' % n)
    return n


class conv_segnet(nn.Module):
    def __init__(self, pretrained=True, disc_arch='resnet', num_critics=6, feedback_seg=False, no_conf=False,
                 modalities='ir_rgb', input_adapter=False, cert_branch=False, arch='custom', late_fusion=False):
        super().__init__()
        if arch != 'pspnet':
            raise NotImplementedError("arch=%r: only arch='pspnet' is on the B200 hot path" % (arch,))
        if feedback_seg or input_adapter or cert_branch:
            raise NotImplementedError("feedback_seg / input_adapter / cert_branch are outside the B200 hot path")
        self.feedback_seg, self.input_adapter, self.no_conf = feedback_seg, input_adapter, no_conf
        self.phase = "train_seg"

        in_channels = _count_input_channels(modalities)
        self.trgb_segnet = build_net.build_network(None, 'resnet50', in_channels=in_channels, late_fusion=late_fusion)
        if late_fusion:
            print('Activated late fusion ...')
        self.trgb_segnet.apply(utils.weights_init_normal)

        if not no_conf:
            taps = _TAP_CHANNELS[bool(late_fusion)][:num_critics]
            print('Creating %d critics....' % len(taps))
            self.critics = nn.ModuleList(create_critic(disc_arch, c) for c in taps)

        if pretrained:
            utils.initModelRenamed(self.trgb_segnet, _PRETRAINED_CHECKPOINT, 'module.', '')

    # ---- phase switch (conf_segnet.py:86-104): requires_grad decides which half of the step builds a graph / gets gradients
    def setLearningModel(self, module, val):
        for p in module.parameters():
            p.requires_grad = val

    def setPhase(self, phase):
        self.phase = phase
        print("Switching to phase: %s" % self.phase)
        if phase not in _PHASES:          # like the reference: an unknown phase only changes the label
            return
        seg_learns, critics_learn = _PHASES[phase]
        if not self.no_conf:
            for critic in self.critics:
                self.setLearningModel(critic, critics_learn)
        self.setLearningModel(self.trgb_segnet, seg_learns)

    # ---- conf_segnet.py:106-140: the same seg net on the day and the night input, one critic per feature tap
    def forward(self, input_a, input_b):
        if PAIR_FORWARD and self.trgb_segnet.pair_ok(input_a, input_b):
            return self._forward_pair(input_a, input_b)
        logits_a, taps_a, cert_a = self.trgb_segnet(*input_a)
        logits_b, taps_b, cert_b = self.trgb_segnet(*input_b)
        output = {}
        if not self.no_conf:
            output['critics_a'] = [critic(taps_a[i]) for i, critic in enumerate(self.critics)]
            output['critics_b'] = [critic(taps_b[i]) for i, critic in enumerate(self.critics)]
        output.update(pred_label_a=logits_a, pred_label_b=logits_b, cert_a=cert_a, cert_b=cert_b, inter_f_b=taps_b)
        return output

    def _forward_pair(self, input_a, input_b):
        """The same dict from ONE seg-net pass over [day batch; night batch] (PSPNet.forward_pair) and ONE pass of every critic over
        the 2B feature taps: outputs are the two halves of the batched results."""
        b = input_a[0].shape[0]
        logits, taps = self.trgb_segnet.forward_pair(input_a, input_b)
        output = {}
        if not self.no_conf:
            maps = [critic(taps[i]) for i, critic in enumerate(self.critics)]
            output['critics_a'] = [m[:b] for m in maps]
            output['critics_b'] = [m[b:] for m in maps]
        output.update(pred_label_a=logits[:b], pred_label_b=logits[b:], cert_a=None, cert_b=None, inter_f_b=[t[b:] for t in taps])
        return output
