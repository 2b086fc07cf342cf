"""Offline "install" of the reference's hot-path modules into baseline/_ref/ (git-ignored, shipped to the GPU box).

The base contract asks for `pip install --target baseline/_ref /root/reference`.  The reference is not a Python package
(no setup.py / pyproject.toml; `environment.yml` is empty), so pip has nothing to build: what such an install would do for a
pure-Python project -- copy its modules -- is done here at file level, for the ~10 files SURVEY.md section 8(a) puts on the
hot path, with their directory layout and contents UNCHANGED (MANIFEST.json records the sha256 of every file).  Nothing of
this is tracked by git or imported by the product (heatnet_pub_b200/); it serves
  * `bench.py --impl reference`: the reference's own PSPNet on the host cores (`cpu_baseline.kind = "reference"`);
  * tests/test_gpu_reference_callers.py: the reference's own conf_segnet.py wrapper executed over the package.
Run in the build container only (`__graft_entry__.build()` calls it when /root/reference exists); on the GPU box the files
that travelled with the snapshot are used as they are.
"""
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HEATNET_REFERENCE", "/root/reference")
DEST = os.path.join(ROOT, "baseline", "_ref")
CM = "models/confusion_maximization"
FILES = [
    f"{CM}/__init__.py", f"{CM}/models/__init__.py", f"{CM}/models/pspnet.py", f"{CM}/models/extractors.py",
    f"{CM}/models/build_net.py", f"{CM}/models/conf_segnet.py", f"{CM}/discriminator_model.py", f"{CM}/utils.py",
    "models/__init__.py", "models/pspnet.py", "models/extractors.py", "models/build_net.py", "scripts/iou_eval.py",
]


def install(force: bool = False) -> bool:
    """-> True when baseline/_ref holds the files afterwards."""
    manifest = os.path.join(DEST, "MANIFEST.json")
    if not os.path.isdir(REF):
        return os.path.exists(manifest)
    if os.path.exists(manifest) and not force:
        return True
    digests = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        digests[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": REF, "files": digests}, open(manifest, "w"), indent=1)
    return True


if __name__ == "__main__":
    print("installed" if install(force=True) else "reference not available")
