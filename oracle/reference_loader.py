"""Import the UNMODIFIED reference modules from baseline/_ref (see baseline/install_reference.py).  TEST / BASELINE
INFRASTRUCTURE ONLY: used by bench.py's reference arm and by tests; never by heatnet_pub_b200/.

The three shims of SURVEY.md section 8(c) are applied around the import, nothing inside the files is touched:
  1. `torchvision.models.resnet.load_state_dict_from_url` (critic_resnet.py:3 expects the pre-0.13 symbol) -- only needed
     when conf_segnet's sibling modules are imported, harmless otherwise;
  2. `build_network` calls `.cuda()` (build_net.py:27): construct `PSPNet(...)` directly on CPU-only hosts;
  3. `pretrained=False` always (no network for model_zoo.load_url).
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
CM_DIR = os.path.join(REF_DIR, "models", "confusion_maximization")


def available() -> bool:
    return os.path.exists(os.path.join(CM_DIR, "models", "pspnet.py"))


def _purge(prefixes):
    for name in [m for m in sys.modules if any(m == p or m.startswith(p + ".") for p in prefixes)]:
        del sys.modules[name]


def load_cm_pspnet():
    """-> the reference's HeatNet `PSPNet` class (models/confusion_maximization/models/pspnet.py), or None."""
    if not available():
        return None
    _purge(["models"])
    sys.path.insert(0, CM_DIR)
    try:
        mod = importlib.import_module("models.pspnet")
        return mod.PSPNet
    finally:
        sys.path.remove(CM_DIR)
        for name in [m for m in sys.modules if m == "models" or m.startswith("models.")]:
            sys.modules["_heatnet_ref_cm." + name] = sys.modules.pop(name)      # keep them alive, free the generic name


def load_iou_eval():
    if not os.path.exists(os.path.join(REF_DIR, "scripts", "iou_eval.py")):
        return None
    spec = importlib.util.spec_from_file_location("_heatnet_ref_iou_eval", os.path.join(REF_DIR, "scripts", "iou_eval.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
