"""The north star's argmax gate (>= 99.9 % agreement, logits within 2e-2 in BF16 / 1e-4 in FP32) ENFORCED on weights that
separate classes: the oracle graph trained in this job by stock torch on class-coloured blocks (oracle/trained_fixture.py,
SURVEY.md appendix D), evaluated at BASELINE config 1's and config 2's frame sizes against the FP32 oracle (torch on the GPU,
TF32 off).  Stock torch's own BF16-autocast agreement on the same frames is printed next to it."""
import pytest
import torch

from oracle import trained_fixture as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trained_sd():
    return T.separated_state_dict("cuda", log=print)


def _net(sd, precision):
    from heatnet_pub_b200 import pspnet
    net = pspnet.PSPNet(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50', in_channels=4,
                        pretrained=False, late_fusion=True)
    net.load_state_dict(sd)
    return net.cuda().eval().set_precision(precision)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("h,w,batch", [(320, 640, 4), (650, 1920, 2)])
def test_argmax_agreement_gate_on_separated_logits(trained_sd, h, w, batch):
    rgb, ir, label = T.eval_frames(batch, h, w, device="cuda")
    ref = torch.cat([T.oracle_logits(trained_sd, rgb[i:i + 1], ir[i:i + 1]) for i in range(batch)])
    floor = torch.cat([T.oracle_logits(trained_sd, rgb[i:i + 1], ir[i:i + 1], autocast_bf16=True) for i in range(batch)])
    rel = lambda a, b: ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()
    if (h, w) == (320, 640):       # the fixture must actually have learnt the task, else the gate is as vacuous as with random weights
        acc = (ref.argmax(1) == label).float().mean().item()
        top2 = ref.topk(2, dim=1).values
        print(f"fixture: eval-mode pixel accuracy {acc:.4f}, median top-2 margin {(top2[:, 0] - top2[:, 1]).median().item():.2f} "
              f"(max |logit| {ref.abs().max().item():.1f})")
        assert acc > 0.97
    for precision, tol, gate in (("bf16", 2e-2, 0.999), ("fp32", 1e-4, 0.9999)):
        net = _net(trained_sd, precision)
        with torch.no_grad():
            logits = net(rgb, ir)[0]
        err, agree = rel(logits, ref), T.agreement(logits, ref)
        print(f"[{precision} {h}x{w} x{batch}, separated-logits fixture] logits rel err {err:.3e}, argmax agreement {agree:.5f} "
              f"(stock torch bf16 autocast on the same frames: rel err {rel(floor, ref):.3e}, agreement {T.agreement(floor, ref):.5f})")
        assert err < tol
        assert agree >= gate
