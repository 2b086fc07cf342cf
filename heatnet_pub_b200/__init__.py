"""heatnet_pub_b200: B200-native (sm_100a) implementation of HeatNet's dense-segmentation hot path.

Public surface = the reference's own module API for this path:
    build_net.build_network, pspnet.PSPNet / PSPModule / PSPUpsample, extractors.resnet50 (+late fusion),
    discriminator_model.FCDiscriminator, conf_segnet.conv_segnet, iou_eval.IoU / ConfusionMatrix.
plus what sits either side of it: losses (fused CrossEntropy / MSE / BCE-with-logits), optim (fused multi-tensor RMSprop / Adam),
parallel (parameter broadcast, bucketed NCCL gradient all-reduce), inputs (the loaders' per-pixel normalisation on the device),
graphs (CUDA-graph replay of a whole training step), utils (weights_init_normal, initModel*, calculate_ious).
All compute runs in libheatnet_b200.so (hand-written CUDA for sm_100a); there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["build_net", "pspnet", "extractors", "discriminator_model", "conf_segnet", "iou_eval", "engine", "losses", "optim", "parallel",
           "inputs", "graphs", "utils"]
