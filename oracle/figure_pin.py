"""TEST INFRASTRUCTURE ONLY -- pins the rasteriser oracle to the only observable output of the real Panda3D renderer that
the reference ships: tests/data/panda3d_obj_batch_render.png / panda3d_obj_scene_render.png.

Those files are the matplotlib figures written by the reference's renderer tests when SAVEFIG is on
(tests/test_batch_renderer_panda3d.py:148-163, tests/test_scene_renderer_panda3d.py): a 640x480-pixel figure with a 2x2
grid of panels showing rgb | normals / depth (cmap gray_r) | binary mask of the scene at
tests/test_batch_renderer_panda3d.py:43-91 (obj_000001 at TWO = quat(.5,.5,-.5,.5), t = (0,0,.3), K fx=fy=300,
c=(320,240), 480x640).  No reference test compares them, but they hold real Panda3D pixels.

This module restates how matplotlib put a 480x640 image into a panel, so that the oracle's render of the same scene can be
compared with the panels quantitatively (tests/test_oracle_figure_pin.py):

  * layout: figure 6.4 x 4.8 in at 100 dpi; default subplot parameters left .125, right .9, bottom .11, top .88,
    wspace = hspace = .2 -> cells of 225.45 x 168 px; imshow(aspect='equal') shrinks the axes to 224 x 168 px centred in
    the cell, i.e. 0.35 figure pixels per image pixel (checked against the black axes rectangles found in the PNG:
    columns 81..305 / 351..575, rows 58..226 / 259..427);
  * resampling: imshow's default interpolation 'antialiased' uses a Hanning window of radius one OUTPUT pixel when an
    image is down-sampled (here by 2.86), applied to the data (scalar panels: before the colour map);
  * colour: rgb / normals panels show the float image as is; depth: Normalize(vmin=min, vmax=max) + gray_r
    (0 = white, max = black); mask: gray.

The sub-pixel position at which Agg snaps the image box is not restated (it depends on the matplotlib version): the
comparison fits one (dx, dy) in [-0.8, 0.8] figure pixels on the MASK panel and uses it for all four panels.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import numpy as np

SCALE = 0.35                              # figure pixels per image pixel
CELL_W = 0.775 / 2.2 * 640                # 225.4545 px
COL_X0 = (80.0 + (CELL_W - 224.0) / 2, 0.547727272727 * 640 + (CELL_W - 224.0) / 2)  # left edge of the image box, columns 0 / 1
ROW_Y0 = (480 - 0.88 * 480, 480 - 0.46 * 480)                                        # top edge of the image box, rows 0 / 1
# figure pixels fully inside the image box and at least one filter radius away from the axes frame
COLS = (np.arange(83, 303), np.arange(354, 573))
ROWS = (np.arange(60, 223), np.arange(262, 425))
PANELS = {"rgb": (0, 0), "normals": (0, 1), "depth": (1, 0), "mask": (1, 1)}  # name -> (row, col)


def crop_panels(figure_rgb: np.ndarray) -> Dict[str, np.ndarray]:
    """The four panels' interior pixels (uint8 [rows, cols, 3]) of a 480x640 reference figure."""
    assert figure_rgb.shape[:2] == (480, 640)
    return {name: np.ascontiguousarray(figure_rgb[np.ix_(ROWS[r], COLS[c])][..., :3]) for name, (r, c) in PANELS.items()}


def _hann_matrix(dest_centres: np.ndarray, n_src: int) -> np.ndarray:
    """[len(dest), n_src] row-normalised Hanning weights; dest_centres are in source-pixel units (pixel k covers [k, k+1])."""
    src_centres = np.arange(n_src) + 0.5
    d = (src_centres[None, :] - dest_centres[:, None]) * SCALE  # distance in OUTPUT pixels
    w = np.where(np.abs(d) < 1.0, 0.5 + 0.5 * np.cos(np.pi * d), 0.0)
    return (w / w.sum(1, keepdims=True)).astype(np.float64)


def resample(img: np.ndarray, row: int, col: int, dx: float = 0.0, dy: float = 0.0) -> np.ndarray:
    """img [C,H,W] -> [len(ROWS[row]), len(COLS[col]), C]: what imshow puts at the panel's interior figure pixels."""
    C, H, W = img.shape
    wx = _hann_matrix((COLS[col] + 0.5 - (COL_X0[col] + dx)) / SCALE, W)
    wy = _hann_matrix((ROWS[row] + 0.5 - (ROW_Y0[row] + dy)) / SCALE, H)
    return np.einsum("yh,chx->yxc", wy, np.einsum("chw,xw->chx", img.astype(np.float64), wx))


def predict_panels(rgb: np.ndarray, normals: np.ndarray, depth: np.ndarray, dx: float = 0.0, dy: float = 0.0) -> Dict[str, np.ndarray]:
    """Figure-space prediction (float, 0..255) of the four panels from a 480x640 render: rgb/normals [3,H,W] in [0,1],
    depth [H,W] metres (0 = background)."""
    zmax = float(depth.max())
    d = resample(depth[None], 1, 0, dx, dy)[..., 0]
    m = resample((depth > 0).astype(np.float64)[None], 1, 1, dx, dy)[..., 0]
    return {
        "rgb": resample(rgb, 0, 0, dx, dy) * 255.0,
        "normals": resample(normals, 0, 1, dx, dy) * 255.0,
        "depth": 255.0 * (1.0 - d / zmax),  # gray_r of Normalize(0, zmax)
        "mask": 255.0 * m,
    }


def fit_offset(depth: np.ndarray, mask_panel: np.ndarray, span: float = 0.8, step: float = 0.1) -> Tuple[float, float, float]:
    """(dx, dy, mean abs error) minimising the mask panel's error: matplotlib/Agg's sub-pixel snapping of the image box."""
    ref = mask_panel.astype(np.float64).mean(-1)
    m = (depth > 0).astype(np.float64)[None]
    best = (0.0, 0.0, np.inf)
    for dx in np.arange(-span, span + 1e-9, step):
        wx = _hann_matrix((COLS[1] + 0.5 - (COL_X0[1] + dx)) / SCALE, m.shape[2])
        mx = np.einsum("chw,xw->chx", m, wx)
        for dy in np.arange(-span, span + 1e-9, step):
            wy = _hann_matrix((ROWS[1] + 0.5 - (ROW_Y0[1] + dy)) / SCALE, m.shape[1])
            e = float(np.abs(255.0 * np.einsum("yh,chx->yxc", wy, mx)[..., 0] - ref).mean())
            if e < best[2]:
                best = (float(dx), float(dy), e)
    return best


def golden_path() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_panda3d_figure_panels.npz")
