"""Generates tests/golden/ref_panda3d_figure_panels.npz from the reference's golden figures (run in the build container,
where /root/reference exists):  the interior pixels of the four panels (rgb, normals, depth, mask) of
tests/data/panda3d_obj_batch_render.png and panda3d_obj_scene_render.png -- the only real Panda3D output in the reference
tree (written by tests/test_batch_renderer_panda3d.py:148-163 / test_scene_renderer_panda3d.py with SAVEFIG).
The panel geometry is stated in oracle/figure_pin.py."""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import figure_pin  # noqa: E402

REF = os.environ.get("HAPPYPOSE_REFERENCE", "/root/reference")


def main():
    out = {}
    for tag, name in (("batch", "panda3d_obj_batch_render.png"), ("scene", "panda3d_obj_scene_render.png")):
        fig = np.asarray(Image.open(os.path.join(REF, "tests", "data", name)).convert("RGB"))
        for k, v in figure_pin.crop_panels(fig).items():
            out[f"{tag}_{k}"] = v
    np.savez_compressed(figure_pin.golden_path(), **out)
    print(figure_pin.golden_path(), {k: v.shape for k, v in out.items()}, os.path.getsize(figure_pin.golden_path()), "bytes")


if __name__ == "__main__":
    main()
