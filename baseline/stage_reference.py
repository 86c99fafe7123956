"""Stages the three reference source files the CPU / GPU reference arms of bench.py execute into the git-ignored
baseline/_ref/ (which travels to the GPU box with the gpurun snapshot; /root/reference does not).

    python baseline/stage_reference.py

Files (unmodified copies, never committed):  pcdet/models/backbones_3d/vfe/{vfe_template,pillar_vfe}.py and
pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py -- they need only torch.  The reference as a package cannot be
installed (pcdet/__init__.py needs a generated version.py, SharedArray, easydict, skimage, spconv ...: SURVEY.md 8c), so
`pip install /root/reference` is not attempted; bench.py imports the staged files by path (baseline/ref_modules.py).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
FILES = ["pcdet/models/backbones_3d/vfe/vfe_template.py", "pcdet/models/backbones_3d/vfe/pillar_vfe.py",
         "pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py"]


def stage() -> bool:
    if not os.path.isdir(REF):
        return False
    dst = os.path.join(HERE, "_ref")
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(dst, os.path.basename(f)))
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged into baseline/_ref/" if ok else "/root/reference not present: nothing staged")
    sys.exit(0 if ok else 1)
