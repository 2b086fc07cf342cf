"""Adversarial / confusion-maximisation wrapper on the B200 kernel library.

Host-side mirror of the reference's `models/confusion_maximization/models/conf_segnet.py:13-140`
(`create_critic`, `conv_segnet` with `.trgb_segnet`, `.critics`, `.setPhase`, `.phase`, same forward dict).
Only the path the north star names is built: arch='pspnet', disc_arch='cyclegan' (FCDiscriminator);
the ResNeXt "custom" arch, ResNet critics, feedback_seg DownNets and the input adapter are out of scope
 and raise NotImplementedError.
"""
import torch
import torch.nn as nn

from . import build_net, discriminator_model, utils
from .utils import weights_init_normal


def create_critic(disc_arch, input_num):
    if disc_arch == 'cyclegan':
        return discriminator_model.FCDiscriminator(input_num)
    raise NotImplementedError("disc_arch=%r: only the default 'cyclegan' critics are on the B200 hot path" % (disc_arch,))


class conv_segnet(nn.Module):
    def __init__(self, pretrained=True, disc_arch='resnet', num_critics=6, feedback_seg=False, no_conf=False,
                 modalities='ir_rgb', input_adapter=False, cert_branch=False, arch='custom', late_fusion=False):
        super(conv_segnet, self).__init__()

        num_input_channels = 0
        if 'rgb' in modalities:
            num_input_channels += 3
            print('Using RGB')
        if 'ir' in modalities:
            num_input_channels += 1
            print('Using IR')
        print('Total numbers of input channels: %d' % (num_input_channels))

        if arch == 'pspnet':
            self.trgb_segnet = build_net.build_network(None, 'resnet50', in_channels=num_input_channels, late_fusion=late_fusion)
            if late_fusion:
                critic_num = [13, 2048, 1024, 512*2, 256*2, 64*2]
                print('Activated late fusion ...')
            else:
                critic_num = [13, 2048, 1024, 512, 256, 64]
        else:
            raise NotImplementedError("arch=%r: only arch='pspnet' is on the B200 hot path" % (arch,))
        if feedback_seg or input_adapter or cert_branch:
            raise NotImplementedError("feedback_seg / input_adapter / cert_branch are outside the B200 hot path")

        self.trgb_segnet.apply(weights_init_normal)
        self.feedback_seg = feedback_seg
        self.input_adapter = input_adapter

        if not no_conf:
            critic_num = critic_num[0:num_critics]
            self.critics = torch.nn.ModuleList()
            print('Creating %d critics....' % (len(critic_num)))
            for i in range(len(critic_num)):
                self.critics.append(create_critic(disc_arch, critic_num[i]))

        if pretrained:
            utils.initModelRenamed(self.trgb_segnet, 'models_finished/training_nc_irrgb_best.pth', 'module.', '')

        self.phase = "train_seg"
        self.no_conf = no_conf

    def setLearningModel(self, module, val):
        for p in module.parameters():
            p.requires_grad = val

    def setPhase(self, phase):
        self.phase = phase
        print("Switching to phase: %s" % self.phase)
        if self.phase == "train_seg":
            if not self.no_conf:
                for c in self.critics:
                    self.setLearningModel(c, False)
            self.setLearningModel(self.trgb_segnet, True)
        elif self.phase == "train_critic":
            if not self.no_conf:
                for c in self.critics:
                    self.setLearningModel(c, True)
            self.setLearningModel(self.trgb_segnet, False)

    def forward(self, input_a, input_b):
        output = {}
        pred_label_day, inter_f_a, cert_a = self.trgb_segnet(*input_a)
        pred_label_night, inter_f_b, cert_b = self.trgb_segnet(*input_b)

        if not self.no_conf:
            output['critics_a'] = []
            output['critics_b'] = []
            for i, c in enumerate(self.critics):
                output['critics_a'].append(c(inter_f_a[i]))
                output['critics_b'].append(c(inter_f_b[i]))

        output['pred_label_a'] = pred_label_day
        output['pred_label_b'] = pred_label_night
        output['cert_a'] = cert_a
        output['cert_b'] = cert_b
        output['inter_f_b'] = inter_f_b
        return output
